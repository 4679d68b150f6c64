timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms'].items()})
PY
