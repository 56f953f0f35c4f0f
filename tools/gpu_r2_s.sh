#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s_bench_m20_n8.json 2> gpurun_out/r2s_bench_m20_n8.err; tail -2 gpurun_out/r2s_bench_m20_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_m20_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','strong','parity_multi_gpu','clocks')})
PY
