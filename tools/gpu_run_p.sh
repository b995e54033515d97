cd $GRAFT_REPO_ROOT
N=${1:-8}
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py > gpurun_out/r2p_ceiling_${N}gpu.json 2> gpurun_out/r2p_ceiling_${N}gpu.err; echo "ceiling rc=$?"; cut -c1-700 gpurun_out/r2p_ceiling_${N}gpu.json
timeout 600 python -m pytest tests -m "gpu and multigpu" -x -q > gpurun_out/r2p_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2p_bench_${N}gpu.json 2> gpurun_out/r2p_bench_${N}gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2p_bench_${N}gpu.err | cut -c1-300
