# gpurun (1 GPU): A/B of two builds of the library on the SAME box (box-to-box spread on this pool is ~5 %, more than most
# kernel changes are worth).  Before the call, build both versions here and park them where the snapshot takes them:
#     python tch-geometric_b200/build.py && cp tch-geometric_b200/tch_geometric/libtchgeo_cuda.so tools/micro/libtchgeo_a.so
#     (change the kernel) ...                                                              ... tools/micro/libtchgeo_b.so
# TCHGEO_LIB makes tch_geometric load that build instead of the in-tree one.
O=gpurun_out/r2misc; mkdir -p $O
for rep in 1 2; do
for v in a b; do
for f in static dynamic; do
  TCHGEO_LIB=$PWD/tools/micro/libtchgeo_$v.so python bench.py --workload temporal --filter $f --steps 5 --warmup 3 --no-cpu > $O/ab.json 2> $O/ab.err
  python -c "
import json; d=json.load(open('$O/ab.json')); print('$v $f: %.3f ms/step' % d['ms_per_step'], [round(x,3) for x in d['roofline']['per_hop_ms']])"
done
done
done
