# gpurun (1 GPU), round 2 job 3: all GPU tests (new: fixed-segment partitioned protocol on virtual ranks), relabel
# persistent kernel after the probing fix (groups x occupancy sweep), partitioned bench on one GPU (fixed vs legacy)
set -x
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -15 $O/gpu_tests.log
for mb in 2 3; do for g in 3 4 6 8; do
  TCHGEO_RELABEL_MINB=$mb TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_m${mb}_g$g.json 2> $O/bench_relabel_m${mb}_g$g.err
  python -c "
import json; d=json.load(open('$O/bench_relabel_m${mb}_g$g.json')); print('minb $mb groups $g: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
done; done
for proto in fixed legacy; do
  timeout 600 python bench.py --workload partitioned --protocol $proto --steps 5 --warmup 3 --no-cpu > $O/bench_part_1gpu_$proto.json 2> $O/bench_part_1gpu_$proto.err
  echo "rc=$?"; tail -2 $O/bench_part_1gpu_$proto.err
  python -c "
import json; d=json.load(open('$O/bench_part_1gpu_$proto.json')); print('$proto: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'], 'e2e', d['e2e'] and d['e2e']['value'])"
done
