# gpurun (1 GPU): the partition form of to_csc / to_csr against the radix form (bit-exact + time), then the CSX tests with it
O=gpurun_out/r2csx; mkdir -p $O
timeout 120 python tools/csx_ab.py > $O/csx_ab.json 2> $O/csx_ab.err; echo "ab rc=$?"; tail -c 1500 $O/csx_ab.json; tail -3 $O/csx_ab.err
TCHGEO_CSX_SORT=partition timeout 120 python -m pytest tests/test_gpu_csx.py tests/test_gpu_fullsize.py tests/test_cpp_harness.py -m gpu -x -q -k "csx or csc or ind2ptr or fixtures or karate_anchor or errors or harness" > $O/gpu_tests_partition.log 2>&1; echo "rc=$?" >> $O/gpu_tests_partition.log; tail -4 $O/gpu_tests_partition.log
