# gpurun --gpus 8, round 2: BASELINE config 5 at its full shape (N = 111 059 956, E = 1 615 685 872 over 8 ranks):
# bit-exact check of the fixed-segment protocol against the replicated sampler, bench lines (fixed, legacy), and the default
# bench line at 8 GPUs (sampling + relabel + walk + hetero)
set -x
O=gpurun_out/r2g; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/check_partitioned.py --scale 1.0 --batches 32 --protocol fixed > $O/check_8gpu_fixed.json 2> $O/check_8gpu_fixed.err; echo "check fixed rc=$?"; cat $O/check_8gpu_fixed.json; tail -3 $O/check_8gpu_fixed.err
timeout 600 $TR bench.py --gpus 8 --workload partitioned --protocol fixed --steps 10 --warmup 3 > $O/bench_part_8gpu_fixed.json 2> $O/bench_part_8gpu_fixed.err
echo "rc=$?"; tail -2 $O/bench_part_8gpu_fixed.err
python -c "
import json; d=json.load(open('$O/bench_part_8gpu_fixed.json')); print('fixed: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'], 'e2e', d['e2e'] and d['e2e']['value'])"
timeout 600 $TR bench.py --gpus 8 --workload partitioned --protocol legacy --steps 5 --warmup 3 --no-e2e > $O/bench_part_8gpu_legacy.json 2> $O/bench_part_8gpu_legacy.err
python -c "
import json; d=json.load(open('$O/bench_part_8gpu_legacy.json')); print('legacy: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'])"
timeout 600 $TR bench.py --gpus 8 --workload partitioned --protocol fixed --slack 1.2 --steps 10 --warmup 3 --no-e2e > $O/bench_part_8gpu_fixed_slack12.json 2> $O/bench_part_8gpu_fixed_slack12.err
python -c "
import json; d=json.load(open('$O/bench_part_8gpu_fixed_slack12.json')); print('fixed slack 1.2: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'])"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_default_8gpu.json 2> $O/bench_default_8gpu.err; echo "default rc=$?"; tail -3 $O/bench_default_8gpu.err
python -c "
import json; d=json.load(open('$O/bench_default_8gpu.json')); print('default 8 GPUs: %.1f G edges/s, e2e %.2f G, walk %.1f G steps/s (%.1f ms), hetero %.1f G edges/s, relabel %.2f ms' % (d['value']/1e9, d['e2e']['value']/1e9, d['walk_steps_per_sec']/1e9, d['walk']['ms_per_step'], d['hetero_edges_per_sec']/1e9, d['with_relabel']['relabel_ms_per_step']))"
