#!/bin/bash
# N-GPU bench of the headline workload under torchrun (gpurun --gpus N -- 'bash tools/gpu/scale_run.sh N')
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_bench_m20_n$N.json 2> gpurun_out/scale_bench_m20_n$N.err
tail -2 gpurun_out/scale_bench_m20_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open('gpurun_out/scale_bench_m20_n$N.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'e2e', 'strong', 'parity_multi_gpu', 'clocks')})
PY
