# per-kernel device times of the fused amortised-guide step at the c3 shape (ncu launch list; compare shares)
export PYTHONPATH=.
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/prof_mvn_bwd.py 2>/dev/null \
  | grep -E "mvn_|mirt_tc" | tail -10 | awk -F'","' '{n=$5; sub(/\(.*/,"",n); v=$NF; gsub(/"/,"",v); printf "%-60s %8.1f us\n", n, v/1000}'
