# gpurun (1 GPU): the three things the driver runs at round end -- GPU tests, smoke(), the default bench line -- and the
# temporal-filter lines after the last kernel change
O=gpurun_out/r2check; mkdir -p $O
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
T0=$(date +%s); python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2check/bench_default_1gpu.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.2f G, %.3f ms, frac %.3f | e2e %s %.3f G (%.1f ms) others %s | relabel %.2f ms | walk %.2f G | hetero %.2f G | cpu %.1f M' % (
  d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], e['transport'], e['value']/1e9, e['ms_per_step'],
  [(o['transport'], round(o['value']/1e9,3)) for o in e.get('other_transports',[])], d['with_relabel']['relabel_ms_per_step'],
  d['walk_steps_per_sec']/1e9, d['hetero_edges_per_sec']/1e9, d['cpu_baseline']['value']/1e6))
PY
for f in static relative dynamic; do python bench.py --workload temporal --filter $f --steps 5 --warmup 3 > $O/bench_temporal_$f.json 2> /dev/null
python -c "
import json; d=json.load(open('$O/bench_temporal_$f.json')); print('temporal $f: %.3f ms/step, %.2f G edges/s, frac %.3f, cpu %.1f M' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['cpu_baseline']['value']/1e6))"
done
