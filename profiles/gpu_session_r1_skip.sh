#!/bin/bash
# validation of idle-window skipping: neighbour + fairness shapes, capped storms, PyTorch tenant
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_fairness.py -x -q -m gpu 2>&1 | tail -n 4
timeout 200 python - <<'PY'
import json, subprocess, sys
sys.path.insert(0, "tests")
import helpers as H
from test_gpu_fairness import gpu0_uuid
res = {}
for name, lib, cap, skip in (("b200_25", H.NEW_SO, 25, "1"), ("b200_25_noskip", H.NEW_SO, 25, "0"), ("b200_10", H.NEW_SO, 10, "1")):
    sb = H.Sandbox()
    extra = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "CUDA_MEM_LIMIT_0": "4g",
             "CUDA_VISIBLE_DEVICES": "0", "LOGGER_LEVEL": "1", "VGPU_B200_SKIP_IDLE_WINDOWS": skip, "CUDA_CORE_LIMIT_0": str(cap)}
    env = H.preload_env(lib, sb, extra, stub=False)
    r = subprocess.run([H.STORM, "--steps", "1000", "--warmup", "1", "--per-step", "200000", "--max-seconds", "12"],
                       env=env, capture_output=True, text=True, timeout=120)
    sb.cleanup()
    d = json.loads(r.stdout.strip().splitlines()[-1]); d.pop("step_wall_s", None); res[name] = d
    print(name, round(d["launches"] / d["wall_s"]), d["p50_ns"], d["p99_ns"], d["gated_launches"], d["sampler_launches"], d["watchdog_loans"], flush=True)
json.dump(res, open("gpurun_out/skip_storm.json", "w"), indent=1)
PY
timeout 200 python -m pytest tests/test_gpu_framework.py -x -q -m gpu -k pytorch_tenant 2>&1 | tail -n 3
cat gpurun_out/neighbour_r1.json; python -c "
import json;f=json.load(open('gpurun_out/fairness_r1.json'))
for k,v in f.items(): print(k,[round(x) for x in v['rates_per_s']])"
