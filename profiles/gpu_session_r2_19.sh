# GPU session 19 of round 2: final confirmation after the synchronous-copy hooks - full GPU suite, then both bench arms
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s19_r2.txt 2>&1; tail -5 gpurun_out/pytest_gpu_s19_r2.txt
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference_r2_s19.json 2> gpurun_out/bench_reference_r2_s19.err; tail -c 600 gpurun_out/bench_reference_r2_s19.json
timeout 400 python bench.py > gpurun_out/bench_r2_s19.json 2> gpurun_out/bench_r2_s19.err; tail -c 1500 gpurun_out/bench_r2_s19.json
