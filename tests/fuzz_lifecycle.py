"""Offline sweep: process lifecycles over one container's shared files.  Tenant 1 runs a random script and is then either
left to exit normally or killed with SIGKILL in mid-life (its ledger record, its lock files and - B200 library - its
footprint-registry entry stay behind); tenant 2 starts in the same sandbox (start-up hygiene: records of dead pids are
purged, loader.c:1580-1673) and runs another random script.  Both transcripts, the ledger bytes left at the end and the
published vgpu.config must equal the compiled reference's.  `python tests/fuzz_lifecycle.py SEED CASES OUT.json`."""
import json
import os
import random
import signal
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H  # noqa: E402
import test_differential_fuzz as F  # noqa: E402


def run_pair(lib, s1, s2, env, kill_first):
    sb = H.Sandbox()
    e = H.preload_env(lib, sb, env)
    p1 = subprocess.Popen([H.SCENARIO], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e)
    if kill_first:
        p1.stdin.write(s1 + "ledger 0\nsleepms 60000\n")
        p1.stdin.flush()
        out1 = []
        n_expected = len(s1.splitlines()) + 1
        while len(out1) < n_expected:
            line = p1.stdout.readline()
            if not line:
                break
            out1.append(line)
        p1.send_signal(signal.SIGKILL)
        p1.wait()
        t1 = "".join(out1)
    else:
        t1, _ = p1.communicate(s1 + "ledger 0\n", timeout=120)
    r2 = subprocess.run([H.SCENARIO], input=s2, capture_output=True, text=True, env=e, timeout=120)
    led = open(sb.ledger(), "rb").read() if os.path.exists(sb.ledger()) else b""
    cfg = sb.config_bytes() if os.path.exists(sb.path("etc/vgpu-manager/config/vgpu.config")) else b""
    sb.cleanup()
    # the two tenants' pids differ between the runs: compare the ledger by its sums, not its pid fields
    sums = []
    for d in range(16):
        base = d * 16392
        n = int.from_bytes(led[base + 16384:base + 16388], "little") if led else 0
        sums.append((n, sorted(int.from_bytes(led[base + 16 * i + 8:base + 16 * i + 16], "little") for i in range(min(n, 1024)))))
    return t1, r2.stdout, r2.returncode, sums, cfg, r2.stderr


def main():
    seed, cases, out = int(sys.argv[1], 0), int(sys.argv[2]), sys.argv[3]
    H.build_all()
    rng = random.Random(seed)
    bad = []
    for case in range(cases):
        env = F.random_env(rng)
        env["VMEMORY_NODE_ENABLED"] = "true"
        if rng.random() < 0.7:
            env["CUDA_MEM_RATIO_0"] = rng.choice(("2", "4"))
        s1 = F.random_script(rng, rng.randrange(5, 30))
        s2 = F.random_script(rng, rng.randrange(5, 30))
        kill_first = rng.random() < 0.6
        if rng.random() < 0.2:
            env["VGPU_B200_SLAB"] = "1"
        ref = run_pair(H.REF_SO, s1, s2, env, kill_first)
        new = run_pair(H.NEW_SO, s1, s2, env, kill_first)
        if ref[:5] != new[:5]:
            bad.append({"case": case, "env": env, "kill_first": kill_first, "s1": s1, "s2": s2, "ref1": ref[0], "new1": new[0], "ref2": ref[1],
                        "new2": new[1], "rc": [ref[2], new[2]], "ledger_ref": str(ref[3][:3]), "ledger_new": str(new[3][:3]),
                        "cfg_equal": ref[4] == new[4], "stderr": new[5][-1500:]})
            with open(out, "w") as f:
                json.dump(bad, f, indent=1)
        if case % 50 == 49:
            print("case", case + 1, "mismatches", len(bad), flush=True)
    print("done: %d cases, %d mismatches" % (cases, len(bad)))


if __name__ == "__main__":
    main()
