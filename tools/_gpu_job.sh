set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
for p in fresh uneven uneven_fresh_i32; do
$TR tools/a2a_bench.py --mb 1024 --pattern $p > gpurun_out/a2a_$p.json 2>/dev/null; tail -1 gpurun_out/a2a_$p.json | cut -c1-160
done
$TR bench.py --gpus 8 --workload partitioned --batches 256 --steps 5 --warmup 2 > gpurun_out/bench_part_8gpu_b256_g1.json 2> gpurun_out/bench_part_8gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_part_8gpu_b256_g1.json')); print('batches 256', d['ms_per_step'], d['value']/1e9, d['phase_ms_per_step_rank0'])"
