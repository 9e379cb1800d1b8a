# gpurun --gpus 4, round 2: the default bench line at 8 GPUs as the driver launches it (sampling + e2e with every host
# transport + relabel + walk + hetero)
O=gpurun_out/r2g4; mkdir -p $O
nproc > $O/host.txt; lscpu | grep -E "Model name|Socket|NUMA node|^CPU\(s\)|Thread" >> $O/host.txt; free -g | head -2 >> $O/host.txt; cat $O/host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
T0=$(date +%s)
timeout 900 $TR bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_default_8gpu.json 2> $O/bench_default_8gpu.err; echo "default rc=$? wall $(( $(date +%s) - T0 )) s"; tail -3 $O/bench_default_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g4/bench_default_8gpu.json').read().strip().splitlines()[-1])
e=d['e2e']
print('4 GPUs: value %.1f G, %.3f ms | e2e %s %.3f G (%.1f ms) others %s | relabel %.2f ms | walk %.2f G (%.2f ms) | hetero %.2f G' % (
  d['value']/1e9, d['ms_per_step'], e['transport'], e['value']/1e9, e['ms_per_step'],
  [(o['transport'], round(o['value']/1e9,3)) for o in e.get('other_transports',[])], d['with_relabel']['relabel_ms_per_step'],
  d['walk_steps_per_sec']/1e9, d['walk']['ms_per_step'], d['hetero_edges_per_sec']/1e9))
PY
