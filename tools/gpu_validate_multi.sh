#!/bin/bash
# N-GPU pass (run under gpurun --gpus N): the multi-GPU tests and the bench line at N ranks.
#   gpurun --gpus 2 --timeout 2400 -- 'bash tools/gpu_validate_multi.sh 2 r2'
N=${1:-2}
TAG=${2:-run}
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -k "multi or peer or shard or devices" > gpurun_out/${TAG}_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_${N}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_${N}gpu.json 2>/dev/null; echo "ref rc=$?"
