cd $GRAFT_REPO_ROOT
out=gpurun_out/r2n_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
run fold_q50; DCT_CUDA_NO_FOLD=1 run nofold_q50
run fold_q95 --quality 95; DCT_CUDA_NO_FOLD=1 run nofold_q95 --quality 95
run fold_q75 --quality 75; run fold_adaptive --adaptive 1; run fold_1080p --W 1920 --H 1080 --frames 256
timeout 120 tools/latency | tee gpurun_out/r2n_latency.jsonl
DCT_CUDA_NO_FOLD=1 timeout 120 tools/latency
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
