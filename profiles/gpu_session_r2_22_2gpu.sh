# GPU session 22 of round 2 (gpurun --gpus 2): the rebalance loop with its table-driven back-off beside the tenants
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_s22.json 2> gpurun_out/bench_n2_s22.err; tail -c 900 gpurun_out/bench_n2_s22.json
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --impl reference > gpurun_out/bench_reference_n2_s22.json 2> gpurun_out/bench_reference_n2_s22.err; tail -c 300 gpurun_out/bench_reference_n2_s22.json
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --no-rebalance > gpurun_out/bench_n2_norebalance_s22.json 2> gpurun_out/bench_n2_norebalance_s22.err; tail -c 300 gpurun_out/bench_n2_norebalance_s22.json
