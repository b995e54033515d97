cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --quality 95 --no-shapes --no-cpu-baseline > gpurun_out/r2r_bench_q95.json 2>/dev/null; echo "q95 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --adaptive 1 --no-shapes --no-cpu-baseline > gpurun_out/r2r_bench_adaptive.json 2>/dev/null; echo "adaptive rc=$?"
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-shapes"
$CMD > gpurun_out/r2r_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2r_launches.csv $CMD > gpurun_out/r2r_ncu_list.log 2>&1; echo "list rc=$?"
$CMD > gpurun_out/r2r_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_fwd_quant_u8_tma|k_dequant_idct_u8_tma|k_replay_fwd_lane|k_replay_inv_lane' -s 12 -c 4 -o gpurun_out/prof_r2r -f $CMD > gpurun_out/r2r_ncu_full.log 2>&1; echo "full rc=$?"
timeout 120 tools/latency | tee gpurun_out/r2r_latency.jsonl
