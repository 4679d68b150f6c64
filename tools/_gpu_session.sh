timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms'].items()})
PY
