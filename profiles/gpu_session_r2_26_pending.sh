# GPU session 26 of round 2 - NOT RUN (the round's 180 GPU-minutes were used up by session 25).  What it would confirm on a
# B200: everything committed after 406b130 that so far is verified on the fake driver only.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
# 1. the random reference trajectories through the device controller / refill kernels (1600 steps)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_reference or controller or refill" > gpurun_out/pytest_gpu_parity_s26_r2.txt 2>&1; tail -3 gpurun_out/pytest_gpu_parity_s26_r2.txt
# 2. on-device readings -> vgpu-smwatcher --source device (writes gpurun_out/smwatcher_device_readings.json: reading vs NVML's own sample)
timeout 300 python -m pytest tests/test_gpu_smwatcher.py -m gpu -q > gpurun_out/pytest_gpu_smwatcher_s26_r2.txt 2>&1; tail -3 gpurun_out/pytest_gpu_smwatcher_s26_r2.txt
# 3. the whole suite and the bench pair on the final build (reference first, like the driver)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s26_r2.txt 2>&1; tail -3 gpurun_out/pytest_gpu_s26_r2.txt
python bench.py --impl reference > gpurun_out/bench_reference_r2_s26.json 2> gpurun_out/bench_reference_r2_s26.err
python bench.py > gpurun_out/bench_r2_s26.json 2> gpurun_out/bench_r2_s26.err
