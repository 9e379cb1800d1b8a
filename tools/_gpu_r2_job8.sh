# gpurun (1 GPU), round 2 job 8: tests (filtered-kernel rewrite, mag-shaped hetero parity, pipelined groups), temporal-filter
# bench lines, weighted sampler on interleaved records, relabel after parallel retry rounds, ncu of the serve kernel
set -x
O=gpurun_out/r2h; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -15 $O/gpu_tests.log
for f in static relative dynamic; do
  timeout 300 python bench.py --workload temporal --filter $f --steps 5 --warmup 3 > $O/bench_temporal_$f.json 2> $O/bench_temporal_$f.err
  python -c "
import json; d=json.load(open('$O/bench_temporal_$f.json')); print('temporal $f: %.3f ms/step, %.1f G edges/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']), d['roofline']['per_hop_ms'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])"
done
timeout 300 python bench.py --sampler weighted --steps 10 --warmup 3 --no-e2e --headline-only > $O/bench_weighted.json 2> $O/bench_weighted.err
python -c "
import json; d=json.load(open('$O/bench_weighted.json')); print('weighted: %.3f ms/step, %.1f G edges/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
timeout 300 python bench.py --sampler replace --steps 10 --warmup 3 --no-e2e --headline-only > $O/bench_replace.json 2> $O/bench_replace.err
python -c "
import json; d=json.load(open('$O/bench_replace.json')); print('replace: %.3f ms/step, %.1f G edges/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
for g in 2 3; do
TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_g$g.json 2> $O/bench_relabel_g$g.err
python -c "
import json; d=json.load(open('$O/bench_relabel_g$g.json')); print('relabel groups $g: %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
done
timeout 300 python bench.py --workload partitioned --protocol fixed --groups 1 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_part_1gpu_g1.json 2> $O/bench_part_1gpu_g1.err
timeout 300 python bench.py --workload partitioned --protocol fixed --groups 2 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_part_1gpu_g2.json 2> $O/bench_part_1gpu_g2.err
for g in 1 2; do python -c "
import json; d=json.load(open('$O/bench_part_1gpu_g$g.json')); print('part 1gpu groups $g: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launch_list_temporal.csv python bench.py --workload temporal --filter static --steps 1 --warmup 1 --no-cpu > $O/ncu_temporal.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hop_filtered -s 5 -c 1 -o $O/r2_hop_filtered python bench.py --workload temporal --filter static --steps 1 --warmup 1 --no-cpu > $O/ncu_filtered.log 2>&1
