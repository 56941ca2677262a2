# GPU session 9 (gpurun --gpus 2): where do the 1.5 % at N = 2 go - rebalance loop or marker spacing?
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 $2 $3 > gpurun_out/$4.log 2> gpurun_out/$4.err; }
run2 29621 --no-rebalance "" bench_n2_norebal
run2 29622 "" "" bench_n2_rebal
run2 29623 --impl reference bench_ref_n2_b
timeout 300 python bench.py --steps 10 --no-extras > gpurun_out/bench_n1_a.log 2> gpurun_out/bench_n1_a.err
timeout 300 python bench.py --steps 10 --no-extras --impl reference > gpurun_out/bench_ref_n1_a.log 2> gpurun_out/bench_ref_n1_a.err
timeout 300 python bench.py --steps 10 --no-extras > gpurun_out/bench_n1_b.log 2> gpurun_out/bench_n1_b.err
for f in bench_n2_norebal bench_n2_rebal bench_ref_n2_b bench_n1_a bench_ref_n1_a bench_n1_b; do echo $f; tail -1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['p50_hook_ns'], d['p99_hook_ns'], d.get('gated_launches'), d.get('rebalance_us'))"; done
