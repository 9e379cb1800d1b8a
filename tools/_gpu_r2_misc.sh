# gpurun (1 GPU): C++ harness with the compact-transport round trip
O=gpurun_out/r2misc; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "harness or gather or transport" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -6 $O/gpu_tests.log
