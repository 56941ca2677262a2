# GPU session 13 of round 2: more runs of the UNMODIFIED reference for the band (its multi-tenant shapes turned out bimodal:
# a fifth gemm4 run gave 0.17-0.21 per tenant where four earlier runs gave 0.215-0.219)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
cp tests/golden/tolerance_band.json gpurun_out/band_acc.json
timeout 900 python tests/band.py --impl reference --runs 8 --shapes gemm4 --append gpurun_out/band_acc.json --out gpurun_out/band_acc.json > gpurun_out/band_more_gemm4.log 2> gpurun_out/band_more_gemm4.err
timeout 600 python tests/band.py --impl reference --runs 4 --shapes gemm1 --append gpurun_out/band_acc.json --out gpurun_out/band_acc.json > gpurun_out/band_more_gemm1.log 2> gpurun_out/band_more_gemm1.err
timeout 900 python tests/band.py --impl reference --runs 4 --shapes storm10,storm25,storm50,neighbour,fair4 --append gpurun_out/band_acc.json --out gpurun_out/band_acc.json > gpurun_out/band_more_cheap.log 2> gpurun_out/band_more_cheap.err
tail -1 gpurun_out/band_more_gemm4.log; tail -1 gpurun_out/band_more_gemm1.log; tail -1 gpurun_out/band_more_cheap.log
