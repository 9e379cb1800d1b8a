# gpurun (1 GPU): host-transfer tests and the e2e number with every transport
O=gpurun_out/r2t; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "gather or transport or harness or partitioned" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
timeout 600 python bench.py --headline-only --no-cpu --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t/bench.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.2f G  e2e %s: %.3f G edges/s, %.2f ms/step, d2h %.2f GB/step @ %.1f GB/s, verified %s' % (d['value']/1e9, e['transport'], e['value']/1e9, e['ms_per_step'], e['d2h_bytes_per_step']/1e9, e['d2h_GBps_per_gpu'], e['landing_zone_verified']))
for o in e.get('other_transports', []): print('  ', o['transport'], '%.3f G, %.2f ms, d2h %.2f GB, verified %s' % (o['value']/1e9, o['ms_per_step'], o['d2h_bytes_per_step']/1e9, o['landing_zone_verified']))
PY
