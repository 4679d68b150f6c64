timeout 900 python -m pytest tests -x -q -m gpu -k "rollout_tc or generation" 2>&1 | tail -4
for r in 1 2; do
for v in "" _head; do
  if [ -z "$v" ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/es_pytorch_b200/libes_b200$v.so; fi
  timeout 300 python tools/dev_tc_time.py 2>&1 | tail -1 | cut -c1-150
done; done
