# A/B of hop-kernel variants inside ONE gpurun call (boxes differ by +-2 %, so variants are only comparable within a call).
# The TPC / DEFER switches it names were experiments of session 3 and no longer exist; TCHGEO_HOP_MIN_BLOCKS does.
set -x
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_partitioned.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/tests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/tests.log; tail -5 $O/tests.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $BARGS > $O/bench_$name.json 2> $O/bench_$name.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); r=d['roofline']
    print('$name', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'serial %.4f'%d['serial']['ms_per_step'], 'frac %.4f'%r['frac'], [round(h['ms'],4) for h in r['per_hop']])
except Exception as e: print('$name FAILED', e)
PY
}
BARGS=""
run tpc2_d0 TCHGEO_HOP_TPC=2 TCHGEO_HOP_DEFER=0
run tpc2_d1 TCHGEO_HOP_TPC=2 TCHGEO_HOP_DEFER=1
run tpc4_d0 TCHGEO_HOP_TPC=4 TCHGEO_HOP_DEFER=0
run tpc4_d1 TCHGEO_HOP_TPC=4 TCHGEO_HOP_DEFER=1
run auto X=1
run tpc2_d0_again TCHGEO_HOP_TPC=2 TCHGEO_HOP_DEFER=0
BARGS="--sampler weighted"
run w_auto X=1
run w_m10 TCHGEO_HOP_MIN_BLOCKS=10
BARGS="--sampler replace"
run r_auto X=1
