# GPU session 21 of round 2: watchdog counts only time the refill thread spends inside the CUDA driver.
# (a) the blocking-copy tenants of session 20 again, (b) the full GPU suite
mkdir -p gpurun_out/loan
python __graft_entry__.py > gpurun_out/build.log 2>&1
sed -i 's/(("dtoh", 4), ("copy", 3), ("sync", 3))/(("dtoh", 4), ("copy", 2), ("htod", 2))/' profiles/gpu_session_r2_20.sh
bash profiles/gpu_session_r2_20.sh 2>&1 | tail -9
cp gpurun_out/loan/summary.json gpurun_out/loan_summary_s21.json
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s21_r2.txt 2>&1; tail -5 gpurun_out/pytest_gpu_s21_r2.txt
