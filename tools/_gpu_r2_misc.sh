# gpurun (1 GPU): negative sampling after the flag kernel went into the scan's input iterator; 32-bit wave tables by default
O=gpurun_out/r2misc; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "negative or relabel or smoke" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
for cfg in "default: " "waves64:TCHGEO_NEG_RELABEL=waves64"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --workload negative --steps 10 --warmup 3 > $O/bench_negative_$name.json 2> $O/bench_negative_$name.err
  python -c "
import json; d=json.load(open('$O/bench_negative_$name.json')); print('$name: %.3f ms/call, %.2f G negatives/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
done
