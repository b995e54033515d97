cd $GRAFT_REPO_ROOT
out=gpurun_out/r2o_kbench.jsonl; : > $out
run() { tag=$1; shift; python tools/kbench.py --tag "$tag" --steps 12 "$@" >> $out 2>&1; tail -1 $out | cut -c1-330; }
run q50; run q95 --quality 95; run one_frame --frames 1 --steps 20; run c3luma --W 7680 --H 4320 --frames 1 --steps 20
timeout 120 tools/latency | tee gpurun_out/r2o_latency.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2o_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_bench_reference.json 2>/dev/null; echo "ref rc=$?"
