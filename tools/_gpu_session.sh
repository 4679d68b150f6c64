timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "rollout_tc or rank_transform or elite or rollout_f32_time" 2>&1 | tail -15 > gpurun_out/r2_memcheck.log; tail -15 gpurun_out/r2_memcheck.log
timeout 600 python bench.py --workload halfcheetah --no-cpu-baseline > gpurun_out/r2_bench_hc.json 2> gpurun_out/r2_bench_hc.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_hc.json').read().strip().splitlines()[-1])
print('halfcheetah', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['config']['pairs_total'], {k:round(v,4) for k,v in d['kernel_ms'].items()})
PY
tail -2 gpurun_out/r2_bench_hc.err
