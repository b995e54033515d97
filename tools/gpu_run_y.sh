cd $GRAFT_REPO_ROOT
./tools/latency > gpurun_out/r2y_latency.jsonl 2> gpurun_out/r2y_latency.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_latency_launches.csv ./tools/latency > gpurun_out/r2y_ncu.log 2>&1; echo "rc=$?"
cat gpurun_out/r2y_latency.jsonl
