# GPU session 24 of round 2: both bench arms at N = 1 with the final build (the driver's own round-end sequence: reference first)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s24.log 2>&1; tail -2 gpurun_out/smoke_s24.log
timeout 200 python bench.py --impl reference > gpurun_out/bench_reference_r2_final.json 2> gpurun_out/bench_reference_r2_final.err; tail -c 200 gpurun_out/bench_reference_r2_final.json
timeout 300 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.json
