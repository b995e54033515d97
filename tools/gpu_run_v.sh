cd $GRAFT_REPO_ROOT
CMD="python tools/kbench.py --tag adaptive --adaptive 1 --steps 3 --frames 64"
$CMD > gpurun_out/r2v_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_fwd_quant_u8_tma|k_dequant_idct_u8_tma' -s 6 -c 2 -o gpurun_out/prof_r2v -f $CMD > gpurun_out/r2v_ncu.log 2>&1; echo "rc=$?"
