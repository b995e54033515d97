cd $GRAFT_REPO_ROOT
out=gpurun_out/r2d_kbench.jsonl; : > $out
for v in 3 8 9 10 11 12 13; do DCT_CUDA_K1_VARIANT=$v DCT_CUDA_K2_VARIANT=$v python tools/kbench.py --tag "variant$v" --steps 15 >> $out 2>&1; tail -1 $out | cut -c1-330; done
for v in 0 3 9 12; do DCT_CUDA_K1_VARIANT=$v DCT_CUDA_K2_VARIANT=$v python tools/kbench.py --tag "1080p_variant$v" --W 1920 --H 1080 --frames 256 --steps 15 >> $out 2>&1; tail -1 $out | cut -c1-330; done
export DCT_CUDA_K1_VARIANT=3 DCT_CUDA_K2_VARIANT=3
python tools/kbench.py --tag v3_q95 --quality 95 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag v3_q90 --quality 90 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag v3_q75 --quality 75 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag v3_zigzag --layout 1 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag v3_adaptive --adaptive 1 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
DCT_CUDA_INV_FP32=1 python tools/kbench.py --tag v3_adaptive_fp32inv --adaptive 1 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag v3_c5strip --W 65536 --H 8192 --frames 1 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
timeout 120 tools/latency | tee gpurun_out/r2d_latency.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
