set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_partitioned.py --scale 0.05 --batches 16 > gpurun_out/check_part_2gpu.json 2> gpurun_out/check_part_2gpu.err; echo "rc=$?"
cat gpurun_out/check_part_2gpu.json
for g in 1 2; do
$TR bench.py --gpus 2 --workload partitioned --batches 256 --steps 5 --warmup 2 --groups $g > gpurun_out/bench_part_2gpu_g$g.json 2> gpurun_out/bench_part_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_part_2gpu_g$g.json')); print('groups $g', d['ms_per_step'], d['value']/1e9, d['phase_ms_per_step_rank0'])"
done
python bench.py --workload partitioned --batches 256 --steps 5 --warmup 2 > gpurun_out/bench_part_1gpu_b256.json 2> gpurun_out/bench_part_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_part_1gpu_b256.json')); print('1gpu', d['ms_per_step'], d['value']/1e9, d['phase_ms_per_step_rank0'])"
