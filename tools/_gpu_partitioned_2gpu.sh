# gpurun --gpus 2: single-GPU partitioned tests, bit-exact check of both exchange protocols, 2-GPU bench of the peer-memory exchange
set -x
O=gpurun_out/r2j; mkdir -p $O
export NCCL_DEBUG=WARN
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q > $O/tests.log 2>&1; echo "rc=$?"; tail -3 $O/tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_partitioned.py --scale 0.05 --batches 16 > $O/check_2gpu.json 2> $O/check_2gpu.err; echo "rc=$?"; cat $O/check_2gpu.json; tail -3 $O/check_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload partitioned --gpus 2 --steps 10 --warmup 3 > $O/bench_part_2gpu_peer.json 2> $O/bench_part_2gpu_peer.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_part_2gpu_peer.json')); print(d['value'], d['ms_per_step'], d['answer_exchange'], d['phase_ms_per_step_rank0'])"
tail -3 $O/bench_part_2gpu_peer.err
