# gpurun (1 GPU): direct-address buckets of the relabel stage -- parity tests, timing against the hashed buckets, per-kernel list
O=gpurun_out/r2n; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "relabel or harness or fullsize or negative" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -6 $O/gpu_tests.log
timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_direct.json 2> $O/bench_relabel_direct.err
python -c "
import json; d=json.load(open('$O/bench_relabel_direct.json')); print('direct: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
TCHGEO_RELABEL_DIRECT=0 timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_hashed.json 2> $O/bench_relabel_hashed.err
python -c "
import json; d=json.load(open('$O/bench_relabel_hashed.json')); print('hashed: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:bk_ -c 12 --csv --log-file $O/launch_list_bk.csv python bench.py --workload relabel --steps 1 --warmup 1 > $O/ncu_bk.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2n/launch_list_bk.csv')) if len(r)>10]
hdr=rows[0]; i={h:k for k,h in enumerate(hdr)}
for r in rows[1:]:
    print(r[i['Kernel Name']][:48], r[i['Metric Name']], r[i['Metric Value']], r[i['Metric Unit']])
PY
for form in direct waves; do
  TCHGEO_NEG_RELABEL=$form python bench.py --workload negative --steps 10 --warmup 3 > $O/bench_negative_$form.json 2> $O/bench_negative_$form.err
  python -c "
import json; d=json.load(open('$O/bench_negative_$form.json')); print('negative, relabel $form: %.3f ms/call, %.2f G negatives/s' % (d['ms_per_step'], d['value']/1e9))"
done
