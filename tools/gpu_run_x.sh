cd $GRAFT_REPO_ROOT
CMD="python tools/kbench.py --tag adaptive --adaptive 1 --steps 3 --frames 64"
$CMD > gpurun_out/r2x_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_replay_inv_lane' -s 3 -c 1 -o gpurun_out/prof_r2x -f $CMD > gpurun_out/r2x_ncu.log 2>&1; echo "rc=$?"
