# gpurun (1 GPU): per-kernel times of one negative-sampling call
O=gpurun_out/r2neg; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"neg_|rl_|bk_|DeviceScan" -c 80 --csv --log-file $O/launch_list.csv python bench.py --workload negative --steps 1 --warmup 1 > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2neg/launch_list.csv')) if len(r)>10]
hdr=rows[0]; i={h:k for k,h in enumerate(hdr)}
cur={}
for r in rows[1:]:
    key=(r[i['ID']], r[i['Kernel Name']][:70])
    cur.setdefault(key,{})[r[i['Metric Name']]]=r[i['Metric Value']]
for (id_,k),m in cur.items():
    print(id_, k, m.get('gpu__time_duration.sum'), m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'))
PY
