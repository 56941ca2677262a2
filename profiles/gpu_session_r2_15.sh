# GPU session 15 of round 2: reference band runs under the settled protocol (2.5 s idle before every run), appended;
# then the B200 library through every shape, twice, and the band tests
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
cp tests/golden/tolerance_band.json gpurun_out/band_acc.json
timeout 900 python tests/band.py --impl reference --runs 4 --shapes storm10,storm25,storm50,neighbour,fair4 --append gpurun_out/band_acc.json --out gpurun_out/band_acc.json > gpurun_out/band_settled_cheap.log 2> gpurun_out/band_settled_cheap.err
timeout 900 python tests/band.py --impl reference --runs 3 --shapes gemm1,gemm4 --append gpurun_out/band_acc.json --out gpurun_out/band_acc.json > gpurun_out/band_settled_gemm.log 2> gpurun_out/band_settled_gemm.err
cp gpurun_out/band_acc.json tests/golden/tolerance_band.json
timeout 1200 python tests/band.py --impl b200 --runs 3 --out gpurun_out/band_b200_s15.json > gpurun_out/band_b200_s15.log 2> gpurun_out/band_b200_s15.err
timeout 900 python -m pytest tests/test_gpu_band.py -m gpu -q --timeout 600 > gpurun_out/pytest_band_s15.log 2>&1
grep -h "\[" gpurun_out/band_settled_gemm.err | cut -c1-200; tail -1 gpurun_out/band_b200_s15.log; tail -4 gpurun_out/pytest_band_s15.log
