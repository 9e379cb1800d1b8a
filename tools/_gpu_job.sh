set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
python bench.py --workload partitioned --batches 64 --steps 5 --warmup 2 > gpurun_out/bench_part_1gpu_b64.json 2> gpurun_out/bench_part_1gpu.err
python bench.py --workload partitioned --batches 256 --steps 5 --warmup 2 > gpurun_out/bench_part_1gpu_b256.json 2>> gpurun_out/bench_part_1gpu.err
cut -c1-200 gpurun_out/bench_part_1gpu_b64.json; cut -c1-200 gpurun_out/bench_part_1gpu_b256.json; tail -3 gpurun_out/bench_part_1gpu.err
