#!/bin/bash
# One GPU-box pass over everything the round's numbers come from (run under gpurun, 1 GPU):
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh r2'
# smoke, the GPU tests, the bench lines (default, reference arm, q95, adaptive), the ncu launch list and one
# `--set full` capture of the four kernels of a step, the single-frame latencies.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-run}
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --quality 95 --no-shapes --no-cpu-baseline > gpurun_out/${TAG}_bench_q95.json 2>/dev/null; echo "q95 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --adaptive 1 --no-shapes --no-cpu-baseline > gpurun_out/${TAG}_bench_adaptive.json 2>/dev/null; echo "adaptive rc=$?"
timeout 120 tools/latency | tee gpurun_out/${TAG}_latency.jsonl
CMD="python bench.py --frames 64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-shapes"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_fwd_quant_u8_tma|k_dequant_idct_u8_tma|k_replay_fwd_lane|k_replay_inv_lane' -s 12 -c 4 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
