# gpurun (1 GPU), round 2 job 1: tests, smoke, the default bench line (now with relabel / walk / hetero), relabel wave sweep,
# launch list + full capture of the relabel kernels, compute-sanitizer over the small tests
set -x
O=gpurun_out/r2a; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -25 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; tail -5 $O/bench_1gpu.err; cut -c1-600 $O/bench_1gpu.json
for mb in 24 48 96; do
  TCHGEO_RELABEL_WAVE_MB=$mb timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_wave$mb.json 2> $O/bench_relabel_wave$mb.err
  python -c "
import json; d=json.load(open('$O/bench_relabel_wave$mb.json')); print('wave $mb MB: relabel %.3f ms, hops %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['hops_ms_per_step'], d['roofline']['frac']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launch_list_relabel.csv python bench.py --workload relabel --steps 1 --warmup 1 --batches 64 > $O/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rl_ -s 30 -c 3 -o $O/r2_relabel python bench.py --workload relabel --steps 1 --warmup 1 --batches 64 > $O/ncu_relabel.log 2>&1
# compute-sanitizer: instrument this library's kernels only (their mangled names contain "tchgeo")
SEL='karate or small_cases or ragged or sampled_trees or serve_kernel or partitioned_plan or peer_ or cumsum_kat or ind2ptr or to_csc_kat or dead_ends or exhausted'
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --kernel-regex kns=tchgeo --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 $O/sanitizer_$tool.log
done
