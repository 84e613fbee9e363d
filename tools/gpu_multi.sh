#!/bin/bash
# usage: tools/gpu_multi.sh N   -- config 5 (train) and config 4 (small_batch) on N GPUs of one box, NCCL comm log kept
N=${1:-2}
mkdir -p gpurun_out
for w in train small_batch; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=gpurun_out/nccl_${w}_n${N}.%h.%p.log \
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_n${N}.log 2> gpurun_out/bench_${w}_n${N}.err
  tail -1 gpurun_out/bench_${w}_n${N}.log | cut -c1-900
done
for f in gpurun_out/nccl_train_n${N}.*.log; do grep -m3 -E "AllReduce|NVLS|Connected all" "$f"; break; done
cat gpurun_out/nccl_train_n${N}.*.log | grep -c "AllReduce" 
rm -f gpurun_out/nccl_small_batch_n${N}.*.log
ls gpurun_out/nccl_train_n${N}.*.log | tail -n +2 | xargs rm -f
