export PYTHONPATH=.
for d in 192 448 256; do
  echo "VIPSY_DBG=$d"
  VIPSY_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:mvn_bwd_dw python scripts/prof_mvn_bwd.py 2>/dev/null | grep mvn_bwd_dw | tail -1 | awk -F'","' '{print $NF}'
done
