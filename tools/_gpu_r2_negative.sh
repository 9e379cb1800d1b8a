# gpurun (1 GPU): negative sampling after keeping every count on the device
O=gpurun_out/r2n; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "negative or relabel or smoke" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --workload negative --steps 10 --warmup 3 > $O/bench_negative.json 2> $O/bench_negative.err
python -c "
import json; d=json.load(open('$O/bench_negative.json')); print('negative: %.3f ms/call, %.2f G negatives/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
