timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for r in 1 2; do
  ES_TC_NO_SHADOW=1 timeout 300 python tools/dev_tc_time.py 2>&1 | tail -1 | cut -c1-150
  timeout 300 python tools/dev_tc_time.py 2>&1 | tail -1 | cut -c1-150
done
