# gpurun (1 GPU), round 2: the measurement job whose outputs are copied into profiles/ (tests, smoke, bench lines, ncu launch
# list + full captures of the dominant kernels)
set -x
O=gpurun_out/r2final; mkdir -p $O
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
T0=$(date +%s); python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"; cut -c1-200 $O/bench_default_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launch_list.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 2 -c 1 -o $O/r2_hop3 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --headline-only > $O/ncu_hop3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bk_ -s 7 -c 7 -o $O/r2_relabel_direct python bench.py --workload relabel --steps 1 --warmup 1 > $O/ncu_relabel.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_ -s 16 -c 4 -o $O/r2_partitioned_hop3 python bench.py --workload partitioned --steps 1 --warmup 2 --no-cpu --no-e2e > $O/ncu_part.log 2>&1
python bench.py --workload walk --steps 5 --warmup 3 > $O/bench_walk.json 2> /dev/null
python bench.py --workload hetero --steps 10 --warmup 3 > $O/bench_hetero.json 2> /dev/null
python bench.py --sampler replace --steps 10 --warmup 3 --no-e2e --headline-only > $O/bench_replace.json 2> /dev/null
python bench.py --sampler weighted --steps 10 --warmup 3 --no-e2e --headline-only > $O/bench_weighted.json 2> /dev/null
for f in static relative dynamic; do python bench.py --workload temporal --filter $f --steps 5 --warmup 3 > $O/bench_temporal_$f.json 2> /dev/null; done
python bench.py --workload negative --steps 10 --warmup 3 > $O/bench_negative.json 2> /dev/null
python bench.py --workload gather --steps 10 --warmup 3 > $O/bench_gather.json 2> /dev/null
python bench.py --workload tempo_walk --steps 5 --warmup 3 > $O/bench_tempo_walk.json 2> /dev/null
python tools/host_unpack_bench.py > $O/host_unpack_bench.json 2> /dev/null
python bench.py --workload relabel --steps 10 --warmup 3 > $O/bench_relabel.json 2> /dev/null
python bench.py --workload partitioned --steps 10 --warmup 3 > $O/bench_partitioned_1gpu.json 2> /dev/null
for f in walk hetero replace weighted temporal_static temporal_relative temporal_dynamic negative gather tempo_walk partitioned_1gpu; do python -c "
import json; d=json.load(open('$O/bench_$f.json')); print('$f', d['metric'], '%.4g'%d['value'], d['unit'], 'ms/step %.3f'%d['ms_per_step'], 'frac', (d.get('roofline') or {}).get('frac'))"; done
