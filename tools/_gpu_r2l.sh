O=gpurun_out/r2l; mkdir -p $O
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $O/bench_sampling_2gpu.json 2> $O/bench_sampling_2gpu.err; echo "rc=$?"
cut -c1-260 $O/bench_sampling_2gpu.json; tail -2 $O/bench_sampling_2gpu.err
