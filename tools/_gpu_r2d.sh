set -x
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_partitioned.py -m gpu -x -q > $O/tests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/tests.log; tail -5 $O/tests.log
run() { TCHGEO_HOP_MIN_BLOCKS=$1 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $2 > $O/bench_$1$3.json 2> $O/bench_$1$3.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$1$3.json')); r=d['roofline']
    print('minb $1 $2', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'serial %.4f'%d['serial']['ms_per_step'], 'frac %.4f'%r['frac'], [round(h['ms'],4) for h in r['per_hop']])
except Exception as e: print('$1 $2 FAILED', e)
PY
}
run 10
run 8
run 10 "--sampler weighted" _weighted
