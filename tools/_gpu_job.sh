set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --workload partitioned --batches 256 --steps 5 --warmup 2 > gpurun_out/bench_part_2gpu_b256.json 2> gpurun_out/bench_part_2gpu.err
python bench.py --workload partitioned --batches 256 --steps 5 --warmup 2 > gpurun_out/bench_part_1gpu_b256.json 2>> gpurun_out/bench_part_1gpu.err
tail -3 gpurun_out/bench_part_2gpu.err
python -c "
import json
for f in ['2gpu','1gpu']:
    d=json.load(open('gpurun_out/bench_part_%s_b256.json'%f)); print(f, d['ms_per_step'], d['value']/1e9, d['phase_ms_per_step_rank0'], d['exchange_bytes_per_step_per_rank'])
"
