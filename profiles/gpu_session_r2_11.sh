# GPU session 11 of round 2: the whole -m gpu suite + smoke on the current build
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_s11.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -15 gpurun_out/pytest_gpu_s11.log; tail -2 gpurun_out/smoke.log
