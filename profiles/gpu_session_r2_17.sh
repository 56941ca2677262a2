# GPU session 17 of round 2: verification of the final build - whole -m gpu suite, smoke, bench both arms (+ 2 repeats)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_final.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_final.log 2> gpurun_out/bench_ref_final.err
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --no-extras --impl reference > gpurun_out/bench_ref_rep$i.log 2>/dev/null
  timeout 300 python bench.py --steps 10 --no-extras > gpurun_out/bench_rep$i.log 2>/dev/null
done
tail -6 gpurun_out/pytest_gpu_final.log; tail -1 gpurun_out/smoke.log; for f in bench_ref_final bench_final bench_ref_rep1 bench_rep1 bench_ref_rep2 bench_rep2; do tail -1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['p50_hook_ns'], d['p99_hook_ns'], d.get('added_p50_ns'), d.get('added_p99_ns'), d.get('achieved_util_pct'), d.get('watchdog_loans'), (d.get('mem_path') or {}).get('counters'))"; done
