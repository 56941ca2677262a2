#!/bin/bash
# Round 1, refill-path session: PyTorch tenant, fairness / neighbour shapes, capped storms
# (default sampler+watchdog mode and VGPU_B200_GOVERNOR=1).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_framework.py -x -q -m gpu > gpurun_out/gov_framework.log 2>&1
echo "framework rc=$?" > gpurun_out/gov_status.txt
timeout 400 python -m pytest tests/test_gpu_fairness.py -x -q -m gpu > gpurun_out/gov_fairness.log 2>&1
echo "fairness rc=$?" >> gpurun_out/gov_status.txt
timeout 200 python - > gpurun_out/gov_storm.log 2>&1 <<'PY'
import json, subprocess, sys
sys.path.insert(0, "tests")
import helpers as H
from test_gpu_fairness import gpu0_uuid
res = {}
for name, lib, cap, gov in (("b200_25", H.NEW_SO, 25, "0"), ("b200_25_governor", H.NEW_SO, 25, "1"),
                            ("ref_25", H.REF_SO, 25, "0"), ("b200_10", H.NEW_SO, 10, "0"), ("ref_10", H.REF_SO, 10, "0")):
    sb = H.Sandbox()
    extra = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "CUDA_MEM_LIMIT_0": "4g",
             "CUDA_VISIBLE_DEVICES": "0", "LOGGER_LEVEL": "1", "VGPU_B200_GOVERNOR": gov}
    if cap: extra["CUDA_CORE_LIMIT_0"] = str(cap)
    env = H.preload_env(lib, sb, extra, stub=False)
    r = subprocess.run([H.STORM, "--steps", "1000", "--warmup", "1", "--per-step", "200000", "--max-seconds", "12"],
                       env=env, capture_output=True, text=True, timeout=120)
    sb.cleanup()
    try: res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception: res[name] = {"rc": r.returncode, "err": r.stderr[-1500:]}
    print(name, json.dumps(res[name])[:600], flush=True)
json.dump(res, open("gpurun_out/gov_storm.json", "w"), indent=1)
PY
echo "storm rc=$?" >> gpurun_out/gov_status.txt
tail -n 5 gpurun_out/gov_framework.log gpurun_out/gov_fairness.log
cat gpurun_out/gov_status.txt
