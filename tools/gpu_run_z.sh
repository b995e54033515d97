cd $GRAFT_REPO_ROOT
python tools/kbench.py --tag new --steps 12 | cut -c1-330
python tools/kbench.py --tag adaptive --adaptive 1 --steps 12 | cut -c1-330
python tools/kbench.py --tag adaptive_q90zz --adaptive 1 --quality 90 --layout 1 --steps 12 | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2z_pytest.log
