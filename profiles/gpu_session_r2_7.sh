# GPU session 7 of round 2: watcher from cuInit + backlog replay - all enforcement shapes again
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
BAND_DETAIL=gpurun_out/diag_fair4_b200_s7.json BAND_LOGGER_LEVEL=3 timeout 400 python tests/band.py --impl b200 --runs 6 --shapes fair4 --out gpurun_out/diag_fair4_band_s7.json > gpurun_out/diag_fair4_s7.log 2>&1
timeout 1200 python tests/band.py --impl b200 --runs 3 --shapes storm10,storm25,storm50,neighbour,gemm1,gemm4 --out gpurun_out/band_b200_s7.json > gpurun_out/band_b200_s7.log 2> gpurun_out/band_b200_s7.err
timeout 900 python -m pytest tests/test_gpu_band.py -m gpu -q --timeout 600 > gpurun_out/pytest_band_s7.log 2>&1
tail -1 gpurun_out/diag_fair4_s7.log; tail -1 gpurun_out/band_b200_s7.log; tail -5 gpurun_out/pytest_band_s7.log
