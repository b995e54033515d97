cd $GRAFT_REPO_ROOT
export DCT_CUDA_K1_VARIANT=3 DCT_CUDA_K2_VARIANT=3
python tools/kbench.py --tag v3_q95 --quality 95 --steps 3 --frames 32 > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_replay_inv_lane|k_replay_fwd_lane' -s 8 -c 2 -o gpurun_out/prof_r2g_k3 -f python tools/kbench.py --tag v3_q95 --quality 95 --steps 3 --frames 32 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2g_plain.log | cut -c1-300
