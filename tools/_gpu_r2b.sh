set -x
O=gpurun_out/r2b; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hop_warp_kernel -s 2 -c 1 -o $O/r2_hop3_warp python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/ncu_hop3.log 2>&1
tail -3 $O/ncu_hop3.log
