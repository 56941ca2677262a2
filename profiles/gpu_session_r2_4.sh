# GPU session 4 of round 2: after the blocking-call fix - fairness diagnostics, the reference band
# regenerated with NUMA-pinned tenants, the B200 library through the same shapes, bench both arms.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
BAND_DETAIL=gpurun_out/diag_fair4_b200_s4.json BAND_LOGGER_LEVEL=3 timeout 300 python tests/band.py --impl b200 --runs 4 --shapes fair4 --out gpurun_out/diag_fair4_band_s4.json > gpurun_out/diag_fair4_s4.log 2>&1
timeout 1500 python tests/band.py --impl reference --runs 6 --shapes storm10,storm25,storm50,neighbour,fair4 --out gpurun_out/tolerance_band_cheap.json > gpurun_out/band_ref_cheap.log 2> gpurun_out/band_ref_cheap.err
timeout 900 python tests/band.py --impl reference --runs 4 --shapes gemm1,gemm4 --out gpurun_out/tolerance_band_gemm.json > gpurun_out/band_ref_gemm.log 2> gpurun_out/band_ref_gemm.err
BAND_DETAIL=gpurun_out/band_b200_detail_s4.json timeout 1200 python tests/band.py --impl b200 --runs 3 --out gpurun_out/band_b200_s4.json > gpurun_out/band_b200_s4.log 2> gpurun_out/band_b200_s4.err
timeout 300 python bench.py --steps 10 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 python -m pytest tests/test_gpu_differential.py tests/test_gpu_slab.py -m gpu -q --timeout 400 > gpurun_out/pytest_gpu_s4.log 2>&1
tail -2 gpurun_out/diag_fair4_s4.log | cut -c1-500; tail -1 gpurun_out/band_ref_cheap.log; tail -1 gpurun_out/band_ref_gemm.log; tail -1 gpurun_out/band_b200_s4.log; tail -1 gpurun_out/bench_ref.log | cut -c1-300; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/pytest_gpu_s4.log
