# gpurun (1 GPU): phase trace of the persistent relabel kernel
O=gpurun_out/r2f; mkdir -p $O
for g in 2 4; do
TCHGEO_RELABEL_TRACE=$O/relabel_trace_g$g.bin TCHGEO_RELABEL_GROUPS=$g timeout 300 python bench.py --workload relabel --steps 1 --warmup 1 > $O/bench_g$g.json 2> $O/bench_g$g.err
python tools/relabel_trace.py $O/relabel_trace_g$g.bin > $O/relabel_trace_g$g.txt; head -45 $O/relabel_trace_g$g.txt
done
