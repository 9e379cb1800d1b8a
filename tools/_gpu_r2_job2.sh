# gpurun (1 GPU), round 2 job 2: persistent relabel kernel -- parity tests, group sweep, ncu capture
set -x
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "relabel or gather or abi" > $O/gpu_tests_relabel.log 2>&1; echo "rc=$?" >> $O/gpu_tests_relabel.log; tail -15 $O/gpu_tests_relabel.log
for g in 2 4 6 8 12; do
  TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_groups$g.json 2> $O/bench_relabel_groups$g.err
  python -c "
import json; d=json.load(open('$O/bench_relabel_groups$g.json')); print('groups $g: relabel %.3f ms, hops %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['hops_ms_per_step'], d['roofline']['frac']))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rl_persistent -s 1 -c 1 -o $O/r2_relabel_persistent python bench.py --workload relabel --steps 1 --warmup 1 > $O/ncu_relabel.log 2>&1
