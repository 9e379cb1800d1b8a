# gpurun (1 GPU): bucketed (shared-memory) relabel form -- parity tests, timing against the persistent form, launch list
O=gpurun_out/r2m; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "relabel or harness or fullsize" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -6 $O/gpu_tests.log
timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_bucketed.json 2> $O/bench_relabel_bucketed.err
python -c "
import json; d=json.load(open('$O/bench_relabel_bucketed.json')); print('bucketed: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
TCHGEO_RELABEL_PERSISTENT=1 TCHGEO_RELABEL_DENSE=1 TCHGEO_RELABEL_GROUPS=4 timeout 300 python bench.py --workload relabel --steps 5 --warmup 3 > $O/bench_relabel_persistent.json 2> $O/bench_relabel_persistent.err
python -c "
import json; d=json.load(open('$O/bench_relabel_persistent.json')); print('persistent: relabel %.3f ms, frac %.3f' % (d['relabel_ms_per_step'], d['roofline']['frac']))"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bk_ -c 12 --csv --log-file $O/launch_list_bk.csv python bench.py --workload relabel --steps 1 --warmup 1 > $O/ncu_bk.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2m/launch_list_bk.csv')) if len(r)>10]
hdr=rows[0]; i={h:k for k,h in enumerate(hdr)}
for r in rows[1:]:
    print(r[i['Kernel Name']][:40], r[i['Metric Name']], r[i['Metric Value']], r[i['Metric Unit']])
PY
