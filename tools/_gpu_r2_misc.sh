# gpurun (1 GPU): A/B of two builds of the temporal-filter kernel on the same box (TCHGEO_LIB)
O=gpurun_out/r2misc; mkdir -p $O
for rep in 1 2; do
for v in b8 b16; do
for f in static dynamic; do
  TCHGEO_LIB=$PWD/tools/micro/libtchgeo_$v.so python bench.py --workload temporal --filter $f --steps 5 --warmup 3 --no-cpu > $O/ab.json 2> $O/ab.err
  python -c "
import json; d=json.load(open('$O/ab.json')); print('$v $f: %.3f ms/step' % d['ms_per_step'], [round(x,3) for x in d['roofline']['per_hop_ms']])"
done
done
done
