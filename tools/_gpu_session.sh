timeout 900 python bench.py > gpurun_out/r1_j_bench.json 2> gpurun_out/r1_j_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1_j_launch_run.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1_j_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline'].get('steps_s'), d['kernel_ms'], d['gpu_launches'], d['clocks'])
PY
