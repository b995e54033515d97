cd $GRAFT_REPO_ROOT
echo "--- 12x1 / 16x1 for small planes"; timeout 120 tools/latency
echo "--- 8x3 for small planes"; DCT_CUDA_SMALL_GEOMETRY=1 timeout 120 tools/latency
python tools/kbench.py --tag adaptive --adaptive 1 --steps 12 | cut -c1-330
timeout 600 python -m pytest tests -m gpu -x -q -k "adaptive or quality or random" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log
