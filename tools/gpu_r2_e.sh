#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
timeout 600 jet_b200/cpp/test_dropin data/_ref/m10.json 2>&1 | tail -3
JET_B200_DEVICES=all timeout 300 jet_b200/cpp/tbc_bench data/_ref/m12.json h5,m,H10,w,y,J,S,G10,P0 --api tbc --reps 2 | tee gpurun_out/r2e_tbc_2dev.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 1 --workload sycamore53_m12_s9 --slices-per-step 16 --lanes 2 > gpurun_out/r2e_bench_m12_n2.json 2> gpurun_out/r2e_bench_m12_n2.err; tail -3 gpurun_out/r2e_bench_m12_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_m12_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','strong','parity_multi_gpu')})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2e_bench_m20_n2.json 2> gpurun_out/r2e_bench_m20_n2.err; tail -3 gpurun_out/r2e_bench_m20_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_m20_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','strong','parity_multi_gpu')})
PY
