# gpurun (1 GPU): A/B of two builds (tools/micro/libtchgeo_a.so, _b.so) on the relabel stage, same box
O=gpurun_out/r2misc; mkdir -p $O
for rep in 1 2; do
for v in a b; do
  TCHGEO_LIB=$PWD/tools/micro/libtchgeo_$v.so python bench.py --workload relabel --steps 5 --warmup 3 > $O/ab.json 2> $O/ab.err
  python -c "
import json; d=json.load(open('$O/ab.json')); print('$v: relabel %.3f ms' % d['relabel_ms_per_step'])"
done
done
TCHGEO_LIB=$PWD/tools/micro/libtchgeo_b.so python -m pytest tests -m gpu -x -q -k "relabel" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -3 $O/gpu_tests.log
