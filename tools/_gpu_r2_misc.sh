# gpurun (1 GPU): tempo_random_walk with aligned 32-byte group loads; compact transport with every sampler
O=gpurun_out/r2misc; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "tempo or transport" > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log; tail -4 $O/gpu_tests.log
python bench.py --workload tempo_walk --steps 5 --warmup 3 --no-cpu > $O/bench_tempo_walk_thread.json 2> $O/bench_tempo_walk_thread.err
python -c "
import json; d=json.load(open('$O/bench_tempo_walk_thread.json')); print('thread: %.3f ms/step, %.3f G/s, frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:tempo_walk -s 1 -c 1 --csv --log-file $O/tempo_walk_thread_ncu.csv python bench.py --workload tempo_walk --steps 1 --warmup 1 --no-cpu > $O/ncu.log 2>&1
grep -E "tempo_walk" $O/tempo_walk_thread_ncu.csv | awk -F'","' '{print $13, $14, $15}' | tail -8
