cd $GRAFT_REPO_ROOT
out=gpurun_out/r2c_kbench.jsonl; : > $out
for v in 0 1 2 3 4 5 6 7; do DCT_CUDA_K1_VARIANT=$v DCT_CUDA_K2_VARIANT=$v python tools/kbench.py --tag "variant$v" --steps 15 >> $out 2>&1; tail -1 $out | cut -c1-330; done
for v in 1 2 4; do DCT_CUDA_K1_VARIANT=$v DCT_CUDA_K2_VARIANT=$v python tools/kbench.py --tag "1080p_variant$v" --W 1920 --H 1080 --frames 256 --steps 15 >> $out 2>&1; tail -1 $out | cut -c1-330; done
python tools/kbench.py --tag q95 --quality 95 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag q90 --quality 90 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
python tools/kbench.py --tag adaptive --adaptive 1 --steps 10 >> $out 2>&1; tail -1 $out | cut -c1-330
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest.log
