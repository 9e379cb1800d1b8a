# gpurun (1 GPU), round 2 job 5: tests after the serve-kernel tuning and the relabel rewrite (split barriers), sweeps
set -x
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -15 $O/gpu_tests.log
for mb in 2 3; do for g in 1 2 3; do
  TCHGEO_RELABEL_MINB=$mb TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_m${mb}_g$g.json 2> $O/bench_relabel_m${mb}_g$g.err
  python -c "
import json; d=json.load(open('$O/bench_relabel_m${mb}_g$g.json')); print('minb $mb groups $g: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
done; done
timeout 600 python bench.py --workload partitioned --protocol fixed --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_part_1gpu_fixed.json 2> $O/bench_part_1gpu_fixed.err
python -c "
import json; d=json.load(open('$O/bench_part_1gpu_fixed.json')); print('fixed: %.3f ms/step, %.1f G edges/s' % (d['ms_per_step'], d['value']/1e9), d['phase_ms_per_step_rank0'])"
TCHGEO_RELABEL_TRACE=$O/relabel_trace.bin TCHGEO_RELABEL_GROUPS=2 timeout 300 python bench.py --workload relabel --steps 1 --warmup 1 > /dev/null 2>&1; python tools/relabel_trace.py $O/relabel_trace.bin | head -60
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pf_serve -s 8 -c 1 -o $O/r2_pf_serve python bench.py --workload partitioned --protocol fixed --steps 1 --warmup 2 --no-cpu --no-e2e > $O/ncu_serve.log 2>&1
