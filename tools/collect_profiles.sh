# copy the outputs of tools/_gpu_r2_final.sh from gpurun_out/r2final into profiles/ under their round-2 names
set -e
F=gpurun_out/r2final; P=profiles
cp $F/gpu_tests.log $P/r2_gpu_tests.log; cp $F/smoke.log $P/r2_smoke.log
for f in bench_default_1gpu bench_reference_arm bench_walk bench_hetero bench_replace bench_weighted bench_temporal_static bench_temporal_relative bench_temporal_dynamic bench_negative bench_gather bench_tempo_walk bench_relabel bench_partitioned_1gpu host_unpack_bench; do cp $F/$f.json $P/r2_$f.json; done
cp $F/launch_list.csv $P/r2_launch_list.csv
cp $F/r2_hop3.ncu-rep $F/r2_relabel_direct.ncu-rep $P/
cp $F/r2_partitioned_hop3.ncu-rep $P/r2_partitioned_hop2.ncu-rep
for r in r2_hop3 r2_relabel_direct r2_partitioned_hop2; do python tools/ncu_summary.py $P/$r.ncu-rep $P/${r}_metrics.csv; done
