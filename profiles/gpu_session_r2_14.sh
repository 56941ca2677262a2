# GPU session 14 of round 2: four 25 % GEMM tenants under the B200 library with per-step logs (hunting the low mode)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
rm -rf $PWD/gpurun_out/gemm4_logs_b200
BAND_STDERR_DIR=$PWD/gpurun_out/gemm4_logs_b200 BAND_LOGGER_LEVEL=4 timeout 400 python tests/band.py --impl b200 --runs 5 --shapes gemm4 --out gpurun_out/x_b200.json > gpurun_out/gemm4_logs_b200.log 2>&1
grep -h "gemm4\[" gpurun_out/gemm4_logs_b200.log; ls gpurun_out/gemm4_logs_b200 | head -30
