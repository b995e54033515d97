cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L | head -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py > gpurun_out/r2i_ceiling_${N}gpu.json 2> gpurun_out/r2i_ceiling_${N}gpu.err; echo "ceiling rc=$?"; cut -c1-900 gpurun_out/r2i_ceiling_${N}gpu.json
python tools/pcie_ceiling.py > gpurun_out/r2i_ceiling_1of${N}gpu.json 2>/dev/null; cut -c1-500 gpurun_out/r2i_ceiling_1of${N}gpu.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r2i_bench_${N}gpu.json 2> gpurun_out/r2i_bench_${N}gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2i_bench_${N}gpu.err
python tools/kbench.py --tag q95 --quality 95 --steps 10 | cut -c1-330
python tools/kbench.py --tag q50 --steps 10 | cut -c1-330
