#!/bin/bash
# 8-GPU box: bench.py at N=8 and N=4 exactly as the driver launches it (one rank per GPU).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/n8_gpus.txt
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$?" >> gpurun_out/n8_status.txt
  tail -n 1 gpurun_out/bench_n$N.log | cut -c1-400
done
cat gpurun_out/n8_status.txt
