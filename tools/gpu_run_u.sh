cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2u_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --adaptive 1 --no-shapes --no-cpu-baseline > gpurun_out/r2u_bench_adaptive.json 2>/dev/null; echo "adaptive rc=$?"
python tools/kbench.py --tag adaptive --adaptive 1 --steps 12 | cut -c1-330
compute-sanitizer --tool memcheck python tools/sanitizer_workload.py > gpurun_out/r2u_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2u_memcheck.log
