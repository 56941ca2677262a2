# GPU session 10 (gpurun --gpus 8): config 5 as BASELINE writes it - 8 tenants, one per B200, 50 % cores each, both arms
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 --impl reference > gpurun_out/bench_ref_n8.log 2> gpurun_out/bench_ref_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
for f in bench_ref_n8 bench_n8; do echo $f; tail -1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['p50_hook_ns'], d['p99_hook_ns'], d.get('gated_launches'), d.get('rebalance'), d.get('watchdog_loans'), d.get('achieved_util_pct'))"; tail -2 gpurun_out/$f.err; done
