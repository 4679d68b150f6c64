timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r1_j_bench_n8.json 2> gpurun_out/r1_j_bench_n8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1_j_bench_n8.json').read().strip().splitlines()[-1])
print('N=8', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_ms'])
PY
