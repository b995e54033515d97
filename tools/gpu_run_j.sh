cd $GRAFT_REPO_ROOT
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-shapes"
$CMD > gpurun_out/r2j_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2j_launches.csv $CMD > gpurun_out/r2j_ncu_list.log 2>&1; echo "list rc=$?"
$CMD > gpurun_out/r2j_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_fwd_quant_u8_tma|k_dequant_idct_u8_tma|k_replay_fwd_lane|k_replay_inv_lane' -s 12 -c 4 -o gpurun_out/prof_r2j -f $CMD > gpurun_out/r2j_ncu_full.log 2>&1; echo "full rc=$?"
tail -1 gpurun_out/r2j_plain.log | cut -c1-200
