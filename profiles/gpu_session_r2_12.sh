# GPU session 12 of round 2: reading / share trajectories of four 25 % GEMM tenants (BASELINE config 3), reference vs B200 library
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
rm -rf gpurun_out/gemm4_logs_ref gpurun_out/gemm4_logs_b200
BAND_STDERR_DIR=gpurun_out/gemm4_logs_ref BAND_LOGGER_LEVEL=4 timeout 300 python tests/band.py --impl reference --runs 2 --shapes gemm4 --out gpurun_out/x_ref.json > gpurun_out/gemm4_logs_ref.log 2>&1
BAND_STDERR_DIR=gpurun_out/gemm4_logs_b200 BAND_LOGGER_LEVEL=4 timeout 300 python tests/band.py --impl b200 --runs 3 --shapes gemm4 --out gpurun_out/x_b200.json > gpurun_out/gemm4_logs_b200.log 2>&1
tail -1 gpurun_out/gemm4_logs_ref.log; tail -1 gpurun_out/gemm4_logs_b200.log; grep -h "gemm4\[" gpurun_out/gemm4_logs_ref.log gpurun_out/gemm4_logs_b200.log
