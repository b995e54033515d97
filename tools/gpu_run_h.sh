cd $GRAFT_REPO_ROOT
out=gpurun_out/r2h_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
run q50; run q95 --quality 95; run q90 --quality 90; run q75 --quality 75; run q10 --quality 10; run zz --layout 1
run adaptive --adaptive 1; run 1080p --W 1920 --H 1080 --frames 256
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
timeout 120 tools/latency | tee gpurun_out/r2h_latency.jsonl
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err
