timeout 600 python -m pytest tests -x -q -m gpu -k "rollout or generation" 2>&1 | tail -5 > gpurun_out/r2_tests.log
timeout 300 python tools/dev_tc_time.py > gpurun_out/r2_time_base.log 2>&1
ES_B200_LIB=$PWD/es_pytorch_b200/libes_b200_trace.so timeout 300 python tools/dev_tc_trace.py > gpurun_out/r2_trace.log 2>&1
tail -3 gpurun_out/r2_tests.log; cat gpurun_out/r2_time_base.log; tail -18 gpurun_out/r2_trace.log
