set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_transform.py tests/test_gpu_sampling.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
python bench.py --sampler weighted --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_weighted2.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_weighted2.json')); r=d['roofline']; print('weighted', d['value']/1e9, d['ms_per_step'], r['frac'], [round(h['ms'],3) for h in r['per_hop']])"
timeout 300 ncu --set full --clock-control none -k regex:gather_rows_kernel -c 1 -o gpurun_out/r1_gather python bench.py --workload gather --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_gather.log 2>&1
