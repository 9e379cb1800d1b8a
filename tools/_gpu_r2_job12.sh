# gpurun (1 GPU): relabel table density x groups sweep; hetero pipelined; real 2-rank pytest is skipped here
O=gpurun_out/r2l; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "partitioned or relabel" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
for d in 0 1; do for g in 2 3 4 6; do
TCHGEO_RELABEL_DENSE=$d TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_d${d}_g$g.json 2> $O/bench_relabel_d${d}_g$g.err
python -c "
import json; d=json.load(open('$O/bench_relabel_d${d}_g$g.json')); print('dense $d groups $g: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
done; done
timeout 300 python bench.py --workload hetero --steps 10 --warmup 3 > $O/bench_hetero.json 2> $O/bench_hetero.err
python -c "
import json; d=json.load(open('$O/bench_hetero.json')); print('hetero: %.3f ms/step pipelined (%.3f serial), %.1f G edges/s, frac %.3f' % (d['ms_per_step'], d['serial_ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
