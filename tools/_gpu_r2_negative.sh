# gpurun (1 GPU): negative sampling -- relabel of the one big tree through direct-address buckets vs the 64-bit wave form
O=gpurun_out/r2neg; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "negative or relabel or smoke" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
for form in direct waves; do
  TCHGEO_NEG_RELABEL=$form python bench.py --workload negative --steps 10 --warmup 3 > $O/bench_negative_$form.json 2> $O/bench_negative_$form.err
  python -c "
import json; d=json.load(open('$O/bench_negative_$form.json')); print('$form: %.3f ms/call, %.2f G negatives/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launch_list.csv python bench.py --workload negative --steps 1 --warmup 1 > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2neg/launch_list.csv')) if len(r)>10]
hdr=rows[0]; i={h:k for k,h in enumerate(hdr)}
seen=[]
for r in rows[1:]:
    seen.append((r[i['Kernel Name']][:60], r[i['Metric Value']]))
for k,v in seen[-30:]: print(k, v)
PY
