# GPU session 2 of round 2: diagnostics of the shapes that missed the reference's band in session 1
# (4-tenant fairness, 50 % storm), allocator A/B, slab mode on the real driver, ncu of the control kernels.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
NEW=vgpu_manager_b200/libvgpu-control.so; REF=oracle/_ref/libvgpu-control.so
# (1) fair4 x3 with every tenant's log
BAND_DETAIL=gpurun_out/diag_fair4_b200.json BAND_LOGGER_LEVEL=3 timeout 300 python tests/band.py --impl b200 --runs 3 --shapes fair4 --out gpurun_out/diag_fair4_band.json > gpurun_out/diag_fair4.log 2>&1
# (2) 50 % storm, both libraries, per-step readings in the log
BAND_DETAIL=gpurun_out/diag_storm50_b200.json BAND_LOGGER_LEVEL=4 timeout 120 python tests/band.py --impl b200 --runs 1 --shapes storm50 --out gpurun_out/diag_storm50_band_b200.json > gpurun_out/diag_storm50_b200.log 2>&1
BAND_DETAIL=gpurun_out/diag_storm50_ref.json BAND_LOGGER_LEVEL=4 timeout 120 python tests/band.py --impl reference --runs 1 --shapes storm50 --out gpurun_out/diag_storm50_band_ref.json > gpurun_out/diag_storm50_ref.log 2>&1
# (3) gemm4 x2 with logs
BAND_DETAIL=gpurun_out/diag_gemm4_b200.json BAND_LOGGER_LEVEL=3 timeout 400 python tests/band.py --impl b200 --runs 2 --shapes gemm4 --out gpurun_out/diag_gemm4_band.json > gpurun_out/diag_gemm4.log 2>&1
# (4) allocator storm, both libraries, twice each (config 4 cap)
timeout 300 python - > gpurun_out/alloc_ab_r2.json 2> gpurun_out/alloc_ab_r2.err <<'PY'
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, helpers as H
H.build_all()
uu = bench.gpu_uuids()
out = {}
for rnd in range(2):
    for name, lib in (("b200", H.NEW_SO), ("reference", H.REF_SO), ("bare", None)):
        out.setdefault(name, []).append(bench.run_allocstorm(H, lib, 0, uu, n=3000))
print(json.dumps(out, indent=1))
PY
# (5) slab mode + client mode on the real driver, and the parity suite
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_differential.py tests/test_gpu_parity.py -m gpu -q --timeout 400 > gpurun_out/pytest_gpu_s2.log 2>&1
# (6) control kernels under ncu again (after the cooperative snapshot)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"vgpu_refill|vgpu_quota|vgpu_vslab" -c 12 -o gpurun_out/ncu_control_r2b python profiles/run_control_kernels.py > gpurun_out/ncu_control_b.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2.log; tail -2 gpurun_out/diag_fair4.log | cut -c1-600; cat gpurun_out/alloc_ab_r2.json | tr -d '\n' | cut -c1-1500
