# gpurun --gpus 2, round 2 job 4: fixed-segment protocol on real ranks -- full-shape-per-rank parity check, bench (fixed vs legacy)
set -x
O=gpurun_out/r2d; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/check_partitioned.py --scale 1.0 --batches 32 --protocol fixed > $O/check_2gpu_fixed.json 2> $O/check_2gpu_fixed.err; echo "check fixed rc=$?"; cat $O/check_2gpu_fixed.json; tail -3 $O/check_2gpu_fixed.err
timeout 600 $TR tools/check_partitioned.py --scale 1.0 --batches 32 --protocol legacy > $O/check_2gpu_legacy.json 2> $O/check_2gpu_legacy.err; echo "check legacy rc=$?"; cat $O/check_2gpu_legacy.json
for proto in fixed legacy; do
  timeout 600 $TR bench.py --gpus 2 --workload partitioned --protocol $proto --steps 5 --warmup 3 > $O/bench_part_2gpu_$proto.json 2> $O/bench_part_2gpu_$proto.err
  echo "rc=$?"; tail -2 $O/bench_part_2gpu_$proto.err
  python -c "
import json; d=json.load(open('$O/bench_part_2gpu_$proto.json')); print('$proto: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'], 'e2e', d['e2e'] and d['e2e']['value'])"
done
timeout 300 $TR -m pytest tests/test_gpu_partitioned.py -q -x -k "fixed" > $O/pytest_2rank.log 2>&1; tail -3 $O/pytest_2rank.log
