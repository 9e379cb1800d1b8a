# gpurun --gpus 2: pipelined batch groups + 16-byte peer stores -- parity at the full per-rank shape, bench groups 1 vs 2
set -x
O=gpurun_out/r2i; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tools/check_partitioned.py --scale 1.0 --batches 32 --protocol fixed --groups 2 > $O/check_2gpu_fixed_g2.json 2> $O/check_2gpu_fixed_g2.err; echo "check rc=$?"; cat $O/check_2gpu_fixed_g2.json; tail -3 $O/check_2gpu_fixed_g2.err
for g in 1 2; do
  timeout 600 $TR bench.py --gpus 2 --workload partitioned --protocol fixed --groups $g --steps 10 --warmup 3 --no-e2e > $O/bench_part_2gpu_g$g.json 2> $O/bench_part_2gpu_g$g.err
  echo "rc=$?"; tail -2 $O/bench_part_2gpu_g$g.err
  python -c "
import json; d=json.load(open('$O/bench_part_2gpu_g$g.json')); print('groups $g: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'])"
done
