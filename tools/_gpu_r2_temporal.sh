# gpurun (1 GPU): temporal-filter tests and bench lines
O=gpurun_out/r2j; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "temporal or filter" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
for f in static relative dynamic; do
  timeout 300 python bench.py --workload temporal --filter $f --steps 5 --warmup 3 --no-cpu > $O/bench_temporal_$f.json 2> $O/bench_temporal_$f.err
  python -c "
import json; d=json.load(open('$O/bench_temporal_$f.json')); print('temporal $f: %.3f ms/step, %.1f G edges/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']), d['roofline']['per_hop_ms'])"
done
