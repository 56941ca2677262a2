"""Offline sweep: the two generators of test_differential_fuzz.py under fresh seeds, many cases, mismatches written to a
file (python tests/fuzz_sweep.py SEED CASES OUT.json).  Not collected by pytest; used to hunt for transcript differences
against the compiled reference on the stub driver before they reach the seeded tests."""
import json
import random
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import helpers as H  # noqa: E402
import test_differential_fuzz as F  # noqa: E402


def main():
    seed, cases, out = int(sys.argv[1], 0), int(sys.argv[2]), sys.argv[3]
    H.build_all()
    rng = random.Random(seed)
    bad = []
    ref_hung = 0
    for case in range(cases):
        script = F.random_script(rng, rng.randrange(8, 60))
        env = F.random_env(rng)
        prep, args = None, ()
        if rng.random() < 0.4:
            env, prep = F.random_membership(rng, env)
        elif rng.random() < 0.3:
            args = ("--gpa",)
        if rng.random() < 0.15:
            env["VGPU_B200_SLAB"] = "1"  # ignored by the reference; accounting must not change
        if prep is None and rng.random() < 0.25:
            # two fake GPUs, listed in either order with an optional hole, each with its own limits; the tenant hops
            # between them (allocations, frees and reports follow the *current* device's host index)
            u0, u1, hole = H.STUB_UUID, "GPU-22222222-2222-2222-2222-222222222222", "GPU-00000000-0000-0000-0000-000000000000"
            order = [u0, u1] if rng.random() < 0.5 else [u1, u0]
            if rng.random() < 0.5:
                order.insert(rng.randrange(3), hole)
            env["MANAGER_VISIBLE_DEVICES"] = ",".join(order)
            env["STUB_GPU_COUNT"] = "2"
            for i in range(len(order)):
                if rng.random() < 0.7:
                    env["CUDA_MEM_LIMIT_%d" % i] = rng.choice(("256m", "1g", "3g"))
                if rng.random() < 0.3:
                    env["CUDA_MEM_RATIO_%d" % i] = rng.choice(("2", "4"))
                if rng.random() < 0.3:
                    env["CUDA_CORE_LIMIT_%d" % i] = rng.choice(("10", "60"))
            env.setdefault("STUB_UTIL", "fixed:5")
            lines = script.splitlines()
            for _ in range(rng.randrange(1, 6)):
                lines.insert(rng.randrange(3, len(lines)), "dev %d" % rng.randrange(2))
            script = "\n".join(lines) + "\n"
        if rng.random() < 0.35:  # entry points and process shapes the seeded generator leaves out
            lines = script.splitlines()
            for _ in range(rng.randrange(1, 5)):
                k = rng.random()
                if k < 0.3:
                    # (no launches in the child under a core limit: the reference does not restart its watcher after a
                    # fork, so a child that outruns the inherited bucket sleeps in rate_limiter for ever)
                    extra = "forkchild %d %d" % (rng.choice((1, 4096, 64 << 20, 1 << 30)), 0 if "CUDA_CORE_LIMIT_0" in env else rng.choice((0, 3, 40)))
                elif k < 0.6:
                    extra = "launchvia %d %d %d %d" % (rng.randrange(9), rng.choice((1, 20)), rng.choice((1, 7, 70000)), rng.choice((1, 3)))
                elif k < 0.8:
                    extra = "graph %d %d\ngraphlaunch %d" % (rng.choice((1, 5)), rng.choice((1, 64)), rng.choice((1, 30)))
                else:
                    extra = rng.choice(("sleepms 120", "totalmem", "nvmlinfo2"))
                lines.insert(rng.randrange(3, len(lines)), extra)
            if rng.random() < 0.3:
                lines.insert(0, "drvver")
            script = "\n".join(lines) + "\n"
            if "CUDA_CORE_LIMIT_0" in env and rng.random() < 0.5:
                env["CUDA_CORE_SOFT_LIMIT_0"] = rng.choice(("60", "100", "5"))
        if rng.random() < 0.15:  # a device reset in the middle of the tenant's life (handles above it go stale in both)
            lines = script.splitlines()
            lines.insert(rng.randrange(3, len(lines)), "reset")
            script = "\n".join(lines) + "\n"
        if rng.random() < 0.12:
            # an NVML-only client (nvidia-smi style, no CUDA context): the report is a host evaluation of the same fold
            # (memgate.c host_nvml_view vs nvml_hook.c:47-103); other tenants' UVA records already in the ledger count
            script = "nvmlinit 0\n" + "".join(rng.choice(("nvmlinfo\n", "nvmlinfo2\n", "persistence\n", "setmode 0\n", "ledger 0\n"))
                                               for _ in range(rng.randrange(2, 7)))
            args = ()
            recs = [(rng.choice((1, 4100000 + rng.randrange(60))), rng.choice((4096, 64 << 20, 3 << 30, (1 << 64) - 1))) for _ in range(rng.randrange(0, 4))]
            inner = prep

            def prep(sb, recs=recs, inner=inner):
                import struct
                if inner:
                    inner(sb)
                if recs:
                    raw = bytearray(262272)
                    for i, (pid, used) in enumerate(recs):
                        struct.pack_into("<iiQ", raw, 16 * i, pid, 0, used)
                    struct.pack_into("<I", raw, 16384, len(recs))
                    with open(sb.ledger(), "wb") as f:
                        f.write(raw)
        try:
            ref = F.run(H.REF_SO, script, env, args, prep)
        except subprocess.TimeoutExpired:
            # the reference never restarts its watcher in a forked child (loader.c fork handling): a child that outruns
            # the bucket it inherited sleeps in rate_limiter for ever.  Nothing to compare against.
            ref_hung += 1
            try:  # the B200 library re-arms its tick thread in the child: the same tenant must get through
                new = F.run(H.NEW_SO, script, env, args, prep)
                print("case", case, "reference hung; b200 library completed rc", new[1], flush=True)
            except subprocess.TimeoutExpired:
                bad.append({"case": case, "env": env, "args": list(args), "script": script, "both_hung": True})
                with open(out, "w") as f:
                    json.dump(bad, f, indent=1)
                print("case", case, "BOTH libraries hung", flush=True)
            continue
        new = F.run(H.NEW_SO, script, env, args, prep)
        if ref[:3] != new[:3]:
            bad.append({"case": case, "env": env, "args": list(args), "script": script, "ref": ref[0], "ref_rc": ref[1], "new": new[0],
                        "new_rc": new[1], "cfg_equal": ref[2] == new[2], "stderr": new[3][-1500:]})
            with open(out, "w") as f:
                json.dump(bad, f, indent=1)
        if case % 100 == 99:
            print("case", case + 1, "mismatches", len(bad), "reference hung", ref_hung, flush=True)
    print("done: %d cases, %d mismatches, reference hung in %d" % (cases, len(bad), ref_hung))


if __name__ == "__main__":
    main()
