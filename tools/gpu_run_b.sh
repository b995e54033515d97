cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
out=gpurun_out/r2b_kbench.jsonl; : > $out
python tools/kbench.py --tag default >> $out 2>&1
DCT_CUDA_NO_TMA=1 python tools/kbench.py --tag no_tma >> $out 2>&1
DCT_CUDA_CTAS_PER_SM=2 python tools/kbench.py --tag ctas2 >> $out 2>&1
DCT_CUDA_CTAS_PER_SM=1 python tools/kbench.py --tag ctas1 >> $out 2>&1
DCT_CUDA_L2PROMO=128 python tools/kbench.py --tag l2promo128 >> $out 2>&1
DCT_CUDA_L2PROMO=0 python tools/kbench.py --tag l2promo0 >> $out 2>&1
python tools/kbench.py --tag zigzag --layout 1 >> $out 2>&1
python tools/kbench.py --tag q95 --quality 95 >> $out 2>&1
python tools/kbench.py --tag q90 --quality 90 >> $out 2>&1
python tools/kbench.py --tag q75 --quality 75 >> $out 2>&1
python tools/kbench.py --tag q10 --quality 10 >> $out 2>&1
python tools/kbench.py --tag adaptive --adaptive 1 >> $out 2>&1
DCT_CUDA_INV_FP32=1 python tools/kbench.py --tag adaptive_fp32inv --adaptive 1 >> $out 2>&1
python tools/kbench.py --tag 1080p --W 1920 --H 1080 --frames 256 >> $out 2>&1
python tools/kbench.py --tag c5strip --W 65536 --H 8192 --frames 1 >> $out 2>&1
cat $out
echo "--- latency (PDL on)"; timeout 120 tools/latency | tee gpurun_out/r2b_latency.jsonl
echo "--- latency (PDL off)"; DCT_CUDA_NO_PDL=1 timeout 120 tools/latency | tee gpurun_out/r2b_latency_nopdl.jsonl
DCT_CUDA_NO_PDL=1 python tools/kbench.py --tag no_pdl >> $out 2>&1; tail -1 $out
