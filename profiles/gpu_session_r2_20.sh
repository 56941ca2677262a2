# GPU session 20 of round 2: where does the single watchdog loan of the dtoh blocking-call test come from?
# 10 % busy tenant, 256 KiB synchronous DtoH every 100 launches, LOGGER_LEVEL=4 (one line per control step, with t=)
mkdir -p gpurun_out/loan
python __graft_entry__.py > gpurun_out/build.log 2>&1
python - <<'P'
import json, os, subprocess, sys
sys.path.insert(0, "tests")
import helpers as H
uuid = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
out = []
for call, rep in (("dtoh", 4), ("copy", 3), ("sync", 3)):
    for k in range(rep):
        sb = H.Sandbox()
        env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": uuid, "CUDA_CORE_LIMIT_0": "10",
                                          "CUDA_MEM_LIMIT_0": "4g", "CUDA_VISIBLE_DEVICES": "0", "LOGGER_LEVEL": "4"}, stub=False)
        r = subprocess.run([H.STORM, "--steps", "1000000", "--warmup", "0", "--per-step", "200", "--max-seconds", "5", "--spin-iters", "20000",
                            "--grid", "592", "--block", "256", "--sync-every", "100", "--block-with", call], env=env, capture_output=True, text=True, timeout=200)
        sb.cleanup()
        d = json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {}
        out.append({"call": call, "run": k, "rc": r.returncode, "launches": d.get("launches"), "gated": d.get("gated_launches"), "loans": d.get("watchdog_loans"),
                    "max_ns": d.get("max_ns")})
        print(out[-1], flush=True)
        if d.get("watchdog_loans"):
            open("gpurun_out/loan/%s_%d.err" % (call, k), "w").write(r.stderr[-400000:])
json.dump(out, open("gpurun_out/loan/summary.json", "w"), indent=1)
P
