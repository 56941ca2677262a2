# GPU session 25 of round 2: the 12 reference trajectories added to tests/golden/watcher.json (limit-range ends, one-SM GPU,
# idle single process, multi-process up_limit ramp with process-count changes) through the device controller and refill kernels
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "controller or refill" > gpurun_out/pytest_gpu_parity_s25_r2.txt 2>&1; tail -4 gpurun_out/pytest_gpu_parity_s25_r2.txt
