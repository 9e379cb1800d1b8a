# gpurun --gpus 8: bit-exact check + bench of the partitioned path (TCHGEO_PEER_ANSWERS=1 forces the peer-memory exchange on 8 ranks)
set -x
O=gpurun_out/r2i; mkdir -p $O
export NCCL_DEBUG=WARN
TCHGEO_PEER_ANSWERS=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_partitioned.py --scale 0.05 --batches 16 > $O/check_8gpu.json 2> $O/check_8gpu.err; echo "rc=$?"; cat $O/check_8gpu.json; tail -3 $O/check_8gpu.err
TCHGEO_PEER_ANSWERS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload partitioned --gpus 8 --steps 10 --warmup 3 > $O/bench_part_8gpu_peer.json 2> $O/bench_part_8gpu_peer.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_part_8gpu_peer.json')); print(d['value'], d['ms_per_step'], d['answer_exchange'], d['phase_ms_per_step_rank0'])"
tail -3 $O/bench_part_8gpu_peer.err
