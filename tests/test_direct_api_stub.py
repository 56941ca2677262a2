"""The direct C ABI (include/vgpu_b200.h PART 2) driven through vgpu_manager_b200.lib on the fake
driver: plumbing only (argument lists, result structs, state hand-over) - the fake GPU executes the
library's kernels through the oracle, so numbers prove nothing about the CUDA code; the GPU twin of
this file is tests/test_gpu_parity.py."""
import json
import os
import subprocess
import sys

import helpers as H

CHILD = r'''
import ctypes as C, json, os, sys
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
import helpers as H
cu = C.CDLL("libcuda.so.1", mode=C.RTLD_GLOBAL)
ctx = C.c_void_p()
dev = C.c_int()
assert cu.cuInit(0) == 0 and cu.cuDeviceGet(C.byref(dev), 0) == 0
assert cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev) == 0 and cu.cuCtxSetCurrent(ctx) == 0
from vgpu_manager_b200 import B200Library
lib = B200Library(env={"MANAGER_VISIBLE_DEVICES": H.STUB_UUID, "MANAGER_COMPATIBILITY_MODE": "0"})
lib.attach()
out = {"version": lib.version()}
# limiter: scripted utilisation through the sampler's tail, then explicit controller steps
lib.limiter_reset(148, 2048, 25, 0, 1, 1)
st = lib.sampler_run(500, 100, 1, 90)
out["sampler"] = [st.share, st.granted - st.consumed, st.steps]
traj = []
for user in (90, 90, 10, 30, 25, 0):
    st = lib.limiter_step(user, user, 1, 1)
    traj.append([st.share, st.granted - st.consumed])
out["steps"] = traj
# default control step: raw samples folded + controller, a few golden steps
import json as _j
tr = _j.load(open(os.path.join(os.environ["REPO"], "tests", "golden", "watcher.json")))["trajectories"][3]
lib.limiter_reset(tr["sm"], tr["thr"], tr["hard"], tr["soft"], tr["core_limit"], tr["hard_limit"])
bucket, rf = 0, []
for i, st in enumerate(tr["steps"][:60]):
    lib.limiter_consume(bucket - st["bucket_in"])
    s = lib.refill(H.util_req_from_golden_step(tr["mode"], st, i + 1))
    rf.append([s.share, s.granted - s.consumed, s.up_limit, s.valid] == st["out"][:4])
    bucket = st["out"][1]
out["refill_ok"] = all(rf)
out["vslab_ops"] = H.vslab_model_check(lib, ops=150)
# memory: quota decision and numbers
q = H.QuotaReq()
q.kind, q.mode, q.n_compute, q.total_memory, q.real_memory = 0, 0, 3, 1 << 30, 1 << 30
for i, b in enumerate((300 << 20, 400 << 20, 200 << 20)):
    q.compute[i].pid, q.compute[i].used_bytes = 100 + i, b
q.request = 200 << 20
res = lib.quota_eval(q, H.QuotaRes())
out["quota"] = [res.used, res.path]
# data movement on the fake device memory
n = (1 << 20) + 48
a, b = C.c_ulonglong(), C.c_ulonglong()
cu.cuMemAlloc_v2.argtypes = [C.POINTER(C.c_ulonglong), C.c_size_t]
assert cu.cuMemAlloc_v2(C.byref(a), n) == 0 and cu.cuMemAlloc_v2(C.byref(b), n) == 0
src = bytes((i * 7 + 3) & 0xFF for i in range(n))
cu.cuMemcpyHtoD_v2.argtypes = [C.c_ulonglong, C.c_char_p, C.c_size_t]
cu.cuMemcpyDtoH_v2.argtypes = [C.c_char_p, C.c_ulonglong, C.c_size_t]
assert cu.cuMemcpyHtoD_v2(a.value, src, n) == 0
lib.spill_copy(b.value, a.value, n, None)
back = C.create_string_buffer(n)
assert cu.cuMemcpyDtoH_v2(back, b.value, n) == 0
out["copy_ok"] = back.raw == src
lib.clear(b.value + 16, n - 32, None)
assert cu.cuMemcpyDtoH_v2(back, b.value, n) == 0
out["clear_ok"] = back.raw[:16] == src[:16] and back.raw[16:n - 16] == bytes(n - 32) and back.raw[n - 16:] == src[n - 16:]
out["slab"] = [lib.slab_insert(0x7000, 4096), lib.slab_remove(0x7000), lib.slab_remove(0x7000)]
print("RESULT " + json.dumps(out))
'''


def test_direct_api_round_trip_on_the_fake_driver():
    import ctypes as C
    H.build_all()
    sb = H.Sandbox()
    env = {k: v for k, v in os.environ.items() if not k.startswith(("CUDA_", "MANAGER_", "VGPU_", "STUB_", "LD_PRELOAD"))}
    env.update({"LD_LIBRARY_PATH": H.STUB_DIR, "REPO": H.ROOT, "VGPU_B200_SANDBOX": sb.dir, "LOGGER_LEVEL": "1"})
    for d in ("tmp/.vgpu_lock", "tmp/.vmem_node"):
        os.makedirs(os.path.join(sb.dir, d), exist_ok=True)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert "sm_100a" in out["version"] and out["copy_ok"] and out["clear_ok"] and out["refill_ok"] and out["vslab_ops"] == 150
    assert out["quota"] == [900 << 20, 2]
    # the same trajectory from the oracle (= reference arithmetic, tests/test_oracle_parity.py)
    o = H.oracle()
    g = H.OrcGpu()
    o.orc_gpu_init(C.byref(g), 148, 2048)
    dev = H.CfgDev(hard_core=25, core_limit=1, hard_limit=1)
    w = H.OrcWatcher()
    o.orc_watcher_init(C.byref(w), C.byref(dev))
    b = C.c_int64(0)
    o.orc_watcher_step(C.byref(g), C.byref(dev), C.byref(w), C.byref(H.OrcUtil(90, 90, 1, 1)), C.byref(b))
    assert out["sampler"][:2] == [w.share, b.value] and out["sampler"][2] == 1
    for user, got in zip((90, 90, 10, 30, 25, 0), out["steps"]):
        o.orc_watcher_step(C.byref(g), C.byref(dev), C.byref(w), C.byref(H.OrcUtil(user, user, 1, 1)), C.byref(b))
        assert got == [w.share, b.value]
