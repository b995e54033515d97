cd $GRAFT_REPO_ROOT
out=gpurun_out/r2s_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
run adaptive --adaptive 1; DCT_CUDA_INV_FP64=1 run adaptive_f64 --adaptive 1; run adaptive_q90zz --adaptive 1 --quality 90 --layout 1; run adaptive_q20 --adaptive 1 --quality 20; run q50
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
