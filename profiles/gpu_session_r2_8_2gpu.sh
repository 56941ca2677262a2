# GPU session 8 of round 2 (gpurun --gpus 2): the N > 1 path of bench.py, both arms, and the 2-GPU tests
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --impl reference > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
timeout 600 python -m pytest tests/test_gpu_differential.py tests/test_gpu_nccl_tenant.py -m gpu -q --timeout 400 > gpurun_out/pytest_gpu_2gpu.log 2>&1
tail -1 gpurun_out/bench_ref_n2.log | cut -c1-600; tail -1 gpurun_out/bench_n2.log | cut -c1-2500; tail -3 gpurun_out/bench_n2.err; tail -3 gpurun_out/pytest_gpu_2gpu.log
