# gpurun (1 GPU): final tree -- GPU tests, smoke(), the two forms of to_csc side by side (with the partition form's per-kernel times)
O=gpurun_out/r2last; mkdir -p $O
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 100 python tools/csx_ab.py > $O/csx_ab.json 2> $O/csx_ab.err; echo "ab rc=$?"; tail -c 1500 $O/csx_ab.json
