cd $GRAFT_REPO_ROOT
python tools/kbench.py --tag adaptive --adaptive 1 --steps 12 | cut -c1-330
python tools/kbench.py --tag adaptive_1080p --adaptive 1 --W 1920 --H 1080 --frames 256 --steps 12 | cut -c1-330
python tools/kbench.py --tag adaptive_q90zz --adaptive 1 --quality 90 --layout 1 --steps 12 | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
