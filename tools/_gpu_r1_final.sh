# gpurun (1 GPU): the measurement job whose outputs are copied into profiles/ (tests, smoke, bench lines, ncu launch list + full capture)
set -x
mkdir -p gpurun_out/final4
O=gpurun_out/final4
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -3 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py > $O/bench_sampling_1gpu.json 2> $O/bench_sampling_1gpu.err
cut -c1-300 $O/bench_sampling_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launch_list.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 2 -c 1 -o $O/r1_hop3_s3 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/ncu_hop3.log 2>&1
python bench.py --workload hetero --steps 10 --warmup 3 > $O/bench_hetero.json 2> /dev/null
python bench.py --sampler replace --steps 10 --warmup 3 --no-e2e > $O/bench_replace.json 2> /dev/null
python bench.py --sampler weighted --steps 10 --warmup 3 --no-e2e > $O/bench_weighted.json 2> /dev/null
for f in hetero replace weighted; do python -c "
import json; d=json.load(open('$O/bench_$f.json')); print('$f', d['metric'], '%.3g'%d['value'], d['unit'], 'ms/step %.3f'%d['ms_per_step'], 'frac', (d.get('roofline') or {}).get('frac'))"; done
