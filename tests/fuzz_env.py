"""Offline sweep for the env -> vgpu.config path (reference loader.c:1927-2052, util.c:27-213): random, deliberately
odd values for every knob the contract names; each case is one `init 0 / totalmem / meminfo` tenant under the compiled
reference and under the B200 library on the fake driver - the 1848 bytes of vgpu.config either library writes and the
transcripts must be identical.  `python tests/fuzz_env.py SEED CASES OUT.json`.  Not collected by pytest."""
import json
import random
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import helpers as H  # noqa: E402
import test_differential_fuzz as F  # noqa: E402

U = ["GPU-11111111-1111-1111-1111-111111111111", "GPU-22222222-2222-2222-2222-222222222222",
     "GPU-00000000-0000-0000-0000-000000000000", "GPU-33333333-3333-3333-3333-333333333333"]
SIZES = ["1g", "1G", "512m", "0.5g", "1.5G", "100", "0", "1k", "4096K", "2t", "1e3m", "1e2", " 2g", "2 g", "0x10m", "-1g", "1gb", "g", "abc",
         "nan", "inf", "1.9999999999g", "18446744073709551615", "99999999999999999999", "1,5g", ".5g", "5.g", "1mm", "07g", "+1g", ""]
CORES = ["10", "25", "100", "0", "-5", "101", "1000", "abc", "7.9", " 30", "3e1", "0x20", "1k", "", "50%", "2147483648"]
RATIOS = ["2", "4", "1", "0.5", "1.0000001", "1.5", "abc", "-2", "1e1", "", "inf", "nan", "0x2", " 3", "2,5"]
BOOLS = ["true", "TRUE", "1", "True", "yes", "false", "0", "", "on", "tRUE", " true", "2"]
MODES = ["0", "1", "2", "100", "101", "102", "200", "3", "99", "150", "201", "300", "-1", "abc", "", "1e2", "0x64"]
TEXT = ["pod-a", "", "x" * 70, "名前", "a b", "ns/with/slash"]


def random_case(rng):
    env = {"LOGGER_LEVEL": "0"}
    if rng.random() < 0.9:
        env["MANAGER_COMPATIBILITY_MODE"] = rng.choice(MODES) if rng.random() < 0.5 else "0"
    shape = rng.random()
    devs = [rng.choice(U) for _ in range(rng.randrange(1, 5))]
    if U[0] not in devs:
        devs[rng.randrange(len(devs))] = U[0]  # the fake GPU the tenant runs on
    if shape < 0.6:
        env["MANAGER_VISIBLE_DEVICES"] = ",".join(devs) + rng.choice(("", ",", ",,"))
    elif shape < 0.8:
        for i, d in enumerate(devs):
            if rng.random() < 0.8:
                env["MANAGER_VISIBLE_DEVICE_%d" % rng.choice((i, i, 15, 16))] = d
    else:
        env["NVIDIA_VISIBLE_DEVICES"] = ",".join(devs)
    for name, vals in (("CUDA_MEM_LIMIT", SIZES), ("CUDA_CORE_LIMIT", CORES), ("CUDA_CORE_SOFT_LIMIT", CORES), ("CUDA_MEM_RATIO", RATIOS),
                       ("CUDA_MEM_OVERSOLD", BOOLS)):
        if rng.random() < 0.5:
            env[name] = rng.choice(vals)
        for i in range(len(devs)):
            if rng.random() < 0.4:
                env["%s_%d" % (name, i)] = rng.choice(vals)
    for name in ("VMEMORY_NODE_ENABLED", "EXTERNAL_SM_WATCHER_ENABLED"):
        if rng.random() < 0.4:
            env[name] = rng.choice(BOOLS)
    for name in ("VGPU_POD_NAME", "VGPU_POD_NAMESPACE", "VGPU_POD_UID", "VGPU_CONTAINER_NAME", "MANAGER_CLIENT_REGISTER_UUID"):
        if rng.random() < 0.3:
            env[name] = rng.choice(TEXT)
    env["STUB_GPU_COUNT"] = "2"
    env["STUB_UTIL"] = "fixed:5"
    return env


def main():
    seed, cases, out = int(sys.argv[1], 0), int(sys.argv[2]), sys.argv[3]
    H.build_all()
    rng = random.Random(seed)
    script = "init 0\ntotalmem\nmeminfo\nnvmlinfo\nalloc 1048576\nmeminfo\n"
    bad, hung = [], 0
    for case in range(cases):
        env = random_case(rng)
        try:
            ref = F.run(H.REF_SO, script, env, ())
        except subprocess.TimeoutExpired:
            hung += 1
            continue
        new = F.run(H.NEW_SO, script, env, ())
        if ref[:3] != new[:3]:
            bad.append({"case": case, "env": env, "ref": ref[0], "ref_rc": ref[1], "new": new[0], "new_rc": new[1],
                        "cfg_equal": ref[2] == new[2], "ref_cfg": ref[2].hex(), "new_cfg": new[2].hex(), "stderr": new[3][-1200:]})
            with open(out, "w") as f:
                json.dump(bad, f, indent=1)
        if case % 100 == 99:
            print("case", case + 1, "mismatches", len(bad), "reference hung", hung, flush=True)
    print("done: %d cases, %d mismatches, reference hung in %d" % (cases, len(bad), hung))


if __name__ == "__main__":
    main()
