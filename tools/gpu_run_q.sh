cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
out=gpurun_out/r2q_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
run q50; run q75 --quality 75; run q90 --quality 90; run q95 --quality 95; run q100 --quality 100; run q10 --quality 10
run adaptive --adaptive 1; DCT_CUDA_INV_FP32=1 run adaptive_fp32inv --adaptive 1; run adaptive_q90 --adaptive 1 --quality 90
