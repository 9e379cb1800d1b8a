# gpurun --gpus 8, round 2: config 5 after the peer rotation + 16-byte peer stores: full-shape parity (1 and 2 batch groups),
# bench with 1 and 2 groups
set -x
O=gpurun_out/r2k; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
for g in 1 2; do
timeout 600 $TR tools/check_partitioned.py --scale 1.0 --batches 32 --protocol fixed --groups $g > $O/check_8gpu_fixed_g$g.json 2> $O/check_8gpu_fixed_g$g.err; echo "check groups $g rc=$?"; cat $O/check_8gpu_fixed_g$g.json
done
for g in 1 2; do
timeout 600 $TR bench.py --gpus 8 --workload partitioned --protocol fixed --groups $g --steps 10 --warmup 3 > $O/bench_part_8gpu_g$g.json 2> $O/bench_part_8gpu_g$g.err
echo "rc=$?"; tail -2 $O/bench_part_8gpu_g$g.err
python -c "
import json; d=json.load(open('$O/bench_part_8gpu_g$g.json')); print('groups $g: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'], 'e2e', d['e2e'] and d['e2e']['value'])"
done
timeout 600 $TR bench.py --gpus 8 --workload partitioned --protocol fixed --groups 4 --steps 10 --warmup 3 --no-e2e > $O/bench_part_8gpu_g4.json 2> $O/bench_part_8gpu_g4.err
python -c "
import json; d=json.load(open('$O/bench_part_8gpu_g4.json')); print('groups 4: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9))"
