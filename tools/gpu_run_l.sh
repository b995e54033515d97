cd $GRAFT_REPO_ROOT
python tools/kbench.py --tag new_adaptive --adaptive 1 --steps 8 | cut -c1-330
cp dct_b200/libdct_cuda.so /tmp/keep.so; cp tools/lib_oldk2f64.so dct_b200/libdct_cuda.so
python tools/kbench.py --tag oldk2_adaptive --adaptive 1 --steps 8 | cut -c1-330
cp /tmp/keep.so dct_b200/libdct_cuda.so
