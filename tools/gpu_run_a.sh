cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
timeout 400 python bench.py --steps 30 --warmup 3 --no-shapes > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
DCT_CUDA_NO_TMA=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-shapes --no-cpu-baseline > gpurun_out/r2a_bench_notma.json 2> gpurun_out/r2a_bench_notma.err; echo "bench notma rc=$?"
timeout 120 tools/tile_skeleton > gpurun_out/r2a_skeleton.jsonl 2>&1; echo "skel rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-shapes --no-cpu-baseline --quality 95 --e2e-steps 2 > gpurun_out/r2a_bench_q95.json 2> gpurun_out/r2a_bench_q95.err; echo "q95 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-shapes --no-cpu-baseline --adaptive 1 > gpurun_out/r2a_bench_adaptive.json 2> gpurun_out/r2a_bench_adaptive.err; echo "adaptive rc=$?"
timeout 200 python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-shapes > gpurun_out/r2a_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_fwd_quant_u8_tma|k_dequant_idct_u8_tma' -s 8 -c 2 -o gpurun_out/prof_r2a -f python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-shapes > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2a_pytest.log
