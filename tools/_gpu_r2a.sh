set -x
O=gpurun_out/r2a; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/tests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/tests.log; tail -15 $O/tests.log
run() { TCHGEO_HOP_KERNEL=$1 TCHGEO_WARP_MIN_BLOCKS=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $3 > $O/bench_$1_$2$4.json 2> $O/bench_$1_$2$4.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$1_$2$4.json')); r=d['roofline']
    print('$1 $2 $3', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'frac %.4f'%r['frac'], [round(h['ms'],4) for h in r['per_hop']])
except Exception as e: print('$1 $2 $3 FAILED', e)
PY
}
run warp 10
if [ $rc -eq 0 ]; then
run cta 10
run warp 8
run warp 12
run warp 10 "--sampler replace" _replace
run warp 10 "--sampler weighted" _weighted
fi
