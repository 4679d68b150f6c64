timeout 600 python bench.py --workload humanoid-nsra --no-cpu-baseline > gpurun_out/r2_bench_nsra.json 2> gpurun_out/r2_bench_nsra.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_nsra.json').read().strip().splitlines()[-1])
print('nsra', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms'].items()})
PY
tail -3 gpurun_out/r2_bench_nsra.err
