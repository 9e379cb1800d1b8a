set -x
O=gpurun_out/r2h; mkdir -p $O
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_partitioned.py --scale 0.05 --batches 16 > $O/check_2gpu.json 2> $O/check_2gpu.err; echo "rc=$?"; cat $O/check_2gpu.json; tail -5 $O/check_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload partitioned --gpus 2 --steps 10 --warmup 3 > $O/bench_part_2gpu_peer.json 2> $O/bench_part_2gpu_peer.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_part_2gpu_peer.json')); print(d['value'], d['ms_per_step'], d['answer_exchange'], d['phase_ms_per_step_rank0'])"
tail -3 $O/bench_part_2gpu_peer.err
