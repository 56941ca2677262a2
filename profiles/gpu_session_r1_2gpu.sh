#!/bin/bash
# 2-GPU box: multi-device differential test + bench at N=2 as the driver launches it.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_differential.py -x -q -m gpu -k "two_real_gpus or reset" 2>&1 | tail -n 15
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "N=2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-300
