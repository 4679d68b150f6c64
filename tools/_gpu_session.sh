for r in 1 2; do
for v in "" _nol1; do
  if [ -z "$v" ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/es_pytorch_b200/libes_b200$v.so; fi
  timeout 300 python tools/dev_tc_time.py 2>&1 | tail -1 | cut -c1-70
done; done
