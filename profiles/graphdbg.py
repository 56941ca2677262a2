import sys, subprocess, os
sys.path.insert(0, "tests")
import helpers as H
import test_gpu_framework as T
for extra in ({"VGPU_B200_GRAPH_LIMIT": "1", "LOGGER_LEVEL": "4"}, {"VGPU_B200_GRAPH_LIMIT": "1", "CUDA_CORE_LIMIT_0": "0"}, {}):
    sb = H.Sandbox()
    knobs = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": T.gpu0_uuid(), "CUDA_VISIBLE_DEVICES": "0",
             "CUDA_MEM_LIMIT_0": "8g", "CUDA_CORE_LIMIT_0": "10", "LOGGER_LEVEL": "2"}
    knobs.update(extra)
    env = H.preload_env(H.NEW_SO, sb, knobs, stub=False)
    r = subprocess.run([sys.executable, "-c", T.GRAPH_TENANT], env=env, capture_output=True, text=True, timeout=200)
    print("=== ours", extra, "rc", r.returncode); print(r.stdout[-300:]); print("\n".join(l for l in r.stderr.splitlines() if ("vGPU" in l and "limiter host" not in l) or "Error" in l)[-6000:])
    sb.cleanup()
sb = H.Sandbox()
env = H.preload_env(H.REF_SO, sb, knobs, stub=False)
r = subprocess.run([sys.executable, "-c", T.GRAPH_TENANT], env=env, capture_output=True, text=True, timeout=200)
print("=== ref rc", r.returncode); print(r.stdout[-300:]); print(r.stderr[-800:])
