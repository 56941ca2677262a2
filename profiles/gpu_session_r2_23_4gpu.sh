# GPU session 23 of round 2 (gpurun --gpus 4): bench at N = 4 with the final build (rebalance back-off at a larger world)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 > gpurun_out/bench_n4_s23.json 2> gpurun_out/bench_n4_s23.err; tail -c 1200 gpurun_out/bench_n4_s23.json
