# GPU session 5 of round 2: after un-serialising the bring-up - fairness shapes again; allocation
# path phase profile; bench both arms with the sparser completion markers.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
BAND_DETAIL=gpurun_out/diag_fair4_b200_s5.json BAND_LOGGER_LEVEL=3 timeout 400 python tests/band.py --impl b200 --runs 6 --shapes fair4 --out gpurun_out/diag_fair4_band_s5.json > gpurun_out/diag_fair4_s5.log 2>&1
BAND_DETAIL=gpurun_out/diag_gemm4_b200_s5.json BAND_LOGGER_LEVEL=3 timeout 600 python tests/band.py --impl b200 --runs 3 --shapes gemm4,neighbour --out gpurun_out/diag_gemm4_band_s5.json > gpurun_out/diag_gemm4_s5.log 2>&1
timeout 300 python - > gpurun_out/alloc_profile_r2.txt 2>&1 <<'PY'
import json, os, subprocess, sys, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, helpers as H
H.build_all()
uu = bench.gpu_uuids()
for name, lib, extra in (("b200", H.NEW_SO, {"VGPU_B200_PROFILE": "1"}), ("b200 (unarmed)", H.NEW_SO, {"VGPU_B200_PROFILE": "1", "VGPU_B200_QUOTA_ARMED": "0"}),
                         ("reference", H.REF_SO, {}), ("b200 again", H.NEW_SO, {"VGPU_B200_PROFILE": "1"}), ("reference again", H.REF_SO, {})):
    ex = {"CUDA_MEM_LIMIT_%d": "8g", "CUDA_MEM_RATIO_%d": "4", "VMEMORY_NODE_ENABLED": "true"}
    ex.update(extra)
    d = bench.run_in_tenant(H, [os.path.join(H.BUILD, "allocstorm"), "--n", "3000", "--bytes", str(1 << 20), "--device", "0"], lib, 0, uu, 0, ex)
    print("==", name, {k: d[k] for k in ("alloc_p50_ns", "alloc_p99_ns", "free_p50_ns", "free_p99_ns", "pairs_per_s")})
    print(d["stderr_tail"])
PY
timeout 300 python bench.py --steps 10 --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --impl reference --no-extras > gpurun_out/bench_ref2.log 2> gpurun_out/bench_ref2.err
timeout 300 python bench.py --steps 10 --no-extras > gpurun_out/bench2.log 2> gpurun_out/bench2.err
tail -1 gpurun_out/diag_fair4_s5.log; tail -1 gpurun_out/diag_gemm4_s5.log; cat gpurun_out/alloc_profile_r2.txt; for f in bench_ref bench bench_ref2 bench2; do tail -1 gpurun_out/$f.log | cut -c1-220; done
