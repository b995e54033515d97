cd $GRAFT_REPO_ROOT
CMD="python tools/kbench.py --tag adaptive --adaptive 1 --steps 2 --frames 32"
$CMD > gpurun_out/r2m_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu.log 2>&1; echo "rc=$?"
grep -E "k_deq|k_replay|k_fwd" gpurun_out/r2m_launches.csv | awk -F'","' '{print $5, $9, $13, $NF}' | sed 's/"//g' | cut -c1-200 | tail -24
