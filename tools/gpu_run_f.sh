cd $GRAFT_REPO_ROOT
out=gpurun_out/r2f_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
for v in 3 10 14 15; do export DCT_CUDA_K1_VARIANT=$v DCT_CUDA_K2_VARIANT=3
  run "k1v${v}_4k"; run "k1v${v}_1080p" --W 1920 --H 1080 --frames 256; run "k1v${v}_c5" --W 65536 --H 8192 --frames 1; run "k1v${v}_8k" --W 7680 --H 4320 --frames 16; run "k1v${v}_zz" --layout 1
done
export DCT_CUDA_K1_VARIANT=3 DCT_CUDA_K2_VARIANT=3
run v3_q95 --quality 95; run v3_q90 --quality 90; run v3_q75 --quality 75; run v3_q10 --quality 10
run v3_adaptive --adaptive 1
DCT_CUDA_INV_FP32=1 run v3_adaptive_fp32inv --adaptive 1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
