#!/bin/bash
# per-kernel durations of one c3 evaluation under the current env
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv python scripts/prof_mirt.py "$@" > /dev/null 2>&1
python scripts/ncu_times.py gpurun_out/l.csv 4
