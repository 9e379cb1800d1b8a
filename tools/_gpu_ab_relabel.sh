# gpurun (1 GPU): relabel tests + the stage's time with direct and hashed buckets
O=gpurun_out/r2misc; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "relabel or negative or harness or fullsize" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -3 $O/gpu_tests.log
python bench.py --workload relabel --steps 10 --warmup 3 > $O/bench_relabel.json 2> $O/ab.err
python -c "
import json; d=json.load(open('$O/bench_relabel.json')); print('direct: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
TCHGEO_RELABEL_DIRECT=0 python bench.py --workload relabel --steps 10 --warmup 3 > $O/bench_relabel_hashed.json 2> $O/ab.err
python -c "
import json; d=json.load(open('$O/bench_relabel_hashed.json')); print('hashed: relabel %.3f ms' % d['relabel_ms_per_step'])"
