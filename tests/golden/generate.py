"""Generate tests/golden/*.json from the REFERENCE's own code (oracle/_ref/ref_cosim).

Run here (the container that has /root/reference) with `python tests/golden/generate.py`.
The vectors are committed; the GPU box and CI only ever read them.  Every case is produced by
executing the reference's functions (delta, change_token, rate_limiter, utilization_watcher,
get_used_gpu_memory_by_device, init_g_vgpu_config_by_env) - nothing here is computed by our
own code, so the files pin the oracle (tests/test_oracle_parity.py) and, through the oracle or
directly, the CUDA kernels (tests/test_gpu_parity.py).
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402

SEED = 0x5EED
B200 = (148, 2048)
GEOMS = [B200, (132, 2048), (108, 2048), (80, 2048), (68, 1536), (16, 1024), (1, 32)]


def gen_delta(co, rng):
    cases = []
    ups = [1, 2, 5, 9, 10, 11, 25, 33, 50, 75, 99, 100]
    for sm, thr in GEOMS:
        total = sm * thr * 32
        for _ in range(60):
            up = rng.choice(ups)
            user = rng.choice([0, 1, up - 1, up, up + 1, up * 2, 100, 150, 400, rng.randint(0, 200)])
            user = max(user, 0)
            share = rng.choice([0, 1, total, total - 1, total // 2, rng.randint(0, total)])
            out = int(co.ask("delta %d %d %d %d %d" % (sm, thr, up, user, share)))
            cases.append([sm, thr, up, user, share, out])
    # degenerate limits the env parser can produce (0 and > 100)
    for up, user, share in [(0, 0, 0), (0, 50, 1000), (200, 10, 5), (1000, 0, 0), (100, 0, 9699328)]:
        out = int(co.ask("delta 148 2048 %d %d %d" % (up, user, share)))
        cases.append([148, 2048, up, user, share, out])
    return cases


def gen_token(co, rng):
    cases = []
    for sm, thr in GEOMS[:4]:
        total = sm * thr * 32
        for _ in range(40):
            bucket = rng.choice([0, -1, -5000, total, total - 1, rng.randint(-total, total)])
            d = rng.choice([0, 1, total, total * 2, rng.randint(0, total), -rng.randint(0, total)])
            out = int(co.ask("token %d %d %d %d" % (sm, thr, bucket, d)))
            cases.append([sm, thr, bucket, d, out])
    return cases


def gen_rate(co, rng):
    cases = []
    dims = [(1, 1, 1), (148, 1, 1), (65535, 65535, 1), (2147483647, 1, 1), (2147483648, 1, 1), (65536, 65536, 1),
            (4294967295, 1, 1), (1024, 1024, 64), (3, 5, 7), (46341, 46341, 1)]
    for gx, gy, gz in dims:
        for bucket in (0, 1, 1000, 9699328):
            out = int(co.ask("rate %d %d %d %d" % (bucket, gx, gy, gz)))
            cases.append([bucket, gx, gy, gz, out])
    return cases


def watcher_traj(name, mode, hard, soft, core_limit, hard_limit, geom, steps, rng, mine_prob=1.0, procs=(1,), util=(0, 100), proc_hold=1):
    """One trajectory of the reference watcher thread, one control step at a time."""
    sb = helpers.Sandbox()
    pids = [4000 + i for i in range(6)]
    mine = {p: (rng.random() < mine_prob) for p in pids}
    mine[pids[0]] = True
    if mode & 2:
        for p in pids:
            d = sb.path("etc/vgpu-manager/.host_proc/%d" % p)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "cgroup"), "w") as f:
                f.write("0::/\n" if mine[p] else "0::/kubepods/other\n")
    co = helpers.Cosim(sb=sb)
    sm, thr = geom
    assert co.ask("winit %d %d %d %d %d %d %d" % (mode, hard, soft, core_limit, hard_limit, sm, thr)) == "ok"
    total = sm * thr * 32
    traj = []
    bucket = 0
    phase_util = rng.randint(*util)
    for t in range(steps):
        # host consumption between control steps (what rate_limiter does to the bucket)
        consume = rng.choice([0, 0, rng.randint(0, total // 8), rng.randint(0, total)])
        bucket_in = bucket - consume
        assert co.ask("wset %d" % bucket_in) == "ok"
        if t % 17 == 0:
            phase_util = rng.randint(*util)
        if t % proc_hold == 0:
            nproc = rng.choice(procs)
        ns = rng.choice([0, 1, 1, 2, 3]) if t % 11 else 0
        samples = []
        for k in range(ns):
            pid = pids[k] if k else pids[0]
            smu = max(0, min(130, phase_util + rng.randint(-10, 10))) if k == 0 else rng.randint(0, 60)
            enc = rng.choice([0, 0, 0, 7, 101])
            dec = rng.choice([0, 0, 3])
            age = rng.choice([0, 0, 0, 100, 999, 1001, 5000])
            samples.append([pid, smu, enc, dec, age, 1 if mine[pid] else 0])
        line = "wstep %d %d %s" % (nproc, ns, " ".join("%d:%d:%d:%d:%d" % tuple(s[:5]) for s in samples))
        share, bucket, up, valid, user, sysu = map(int, co.ask(line).split())
        traj.append({"bucket_in": bucket_in, "nproc": nproc, "samples": samples,
                     "out": [share, bucket, up, valid, user, sysu]})
    co.close()
    return {"name": name, "mode": mode, "hard": hard, "soft": soft, "core_limit": core_limit,
            "hard_limit": hard_limit, "sm": sm, "thr": thr, "steps": traj}


def gen_watcher(rng):
    out = []
    out.append(watcher_traj("hard25_host", 0, 25, 0, 1, 1, B200, 160, rng))
    out.append(watcher_traj("hard10_host", 0, 10, 0, 1, 1, B200, 120, rng))
    out.append(watcher_traj("hard50_small_gpu", 0, 50, 0, 1, 1, (16, 1024), 100, rng))
    out.append(watcher_traj("balance25_60_cgv2", 2, 25, 60, 1, 0, B200, 200, rng, mine_prob=0.5, procs=(1, 2, 2, 3)))
    out.append(watcher_traj("balance50_100_host", 0, 50, 100, 1, 0, B200, 160, rng, procs=(1, 1, 2, 4)))
    out.append(watcher_traj("hard25_cgv2_multi", 2, 25, 0, 1, 1, B200, 120, rng, mine_prob=0.4, procs=(1, 2, 3)))
    out.append(watcher_traj("nolimit", 0, 0, 0, 0, 0, B200, 20, rng))
    # round 2 (appended, so that the vectors above keep their random stream): the ends of the limit range, a
    # GPU whose whole bucket is 1024 tokens, a nearly idle single process (the "write the bucket directly"
    # guard, cuda_hook.c:424-427), long balance-mode runs through several up_limit ramps and process changes
    out.append(watcher_traj("hard100_host", 0, 100, 0, 1, 1, B200, 100, rng))
    out.append(watcher_traj("hard1_host", 0, 1, 0, 1, 1, B200, 100, rng, util=(0, 12)))
    out.append(watcher_traj("hard5_one_sm", 0, 5, 0, 1, 1, (1, 32), 100, rng))
    out.append(watcher_traj("hard25_idle_single", 0, 25, 0, 1, 1, B200, 160, rng, util=(0, 4)))
    out.append(watcher_traj("hard40_h100_saturated", 0, 40, 0, 1, 1, (132, 2048), 120, rng, util=(90, 130)))
    out.append(watcher_traj("balance10_30_host_long", 0, 10, 30, 1, 0, B200, 320, rng, procs=(1, 1, 1, 2, 3), util=(0, 40)))
    out.append(watcher_traj("balance75_100_cgv2", 2, 75, 100, 1, 0, B200, 240, rng, mine_prob=0.6, procs=(1, 2), util=(20, 100)))
    out.append(watcher_traj("balance30_90_small_gpu_idle_system", 0, 30, 90, 1, 0, (68, 1536), 320, rng, util=(0, 15)))
    # the multi-process up_limit ramp (cuda_hook.c:451-462) needs a process count that stays >= 2 for 30 ticks
    out.append(watcher_traj("balance50_80_two_procs_ramp", 0, 50, 80, 1, 0, B200, 280, rng, procs=(2,), util=(0, 20)))
    out.append(watcher_traj("balance20_100_procs_come_and_go", 0, 20, 100, 1, 0, B200, 420, rng, procs=(2, 3, 4), util=(0, 30), proc_hold=97))
    out.append(watcher_traj("balance20_60_three_procs_busy_system", 2, 20, 60, 1, 0, B200, 200, rng, mine_prob=0.5, procs=(3,), util=(80, 120)))
    out.append(watcher_traj("balance5_50_two_procs_zero_step", 0, 5, 50, 1, 0, B200, 100, rng, procs=(2,), util=(0, 10)))
    return out


RANDOM_SEED = 0xB200
RANDOM_TRAJECTORIES = 16


def random_trajectory(rng, steps, name):
    """Parameters drawn at random (limits, soft limit on either side of the hard one, geometry, process counts, host /
    cgroup-v2 membership - the two modes watcher_traj can prepare membership files for)."""
    mode = rng.choice([0, 2])
    hard = rng.choice([1, 2, 5, 10, 20, 25, 33, 50, 75, 99, 100])
    balance = rng.random() < 0.45
    soft = rng.choice([hard + 1, min(100, hard + rng.randint(1, 60)), 100]) if balance else rng.choice([0, hard, max(0, hard - 5)])
    lo = rng.choice([0, 0, 10, 40, 90])
    util = (lo, lo + rng.choice([5, 20, 60, 110]))
    procs = rng.choice([(1,), (1, 2), (2,), (1, 1, 2, 4), (3,), (2, 3, 4)])
    return watcher_traj(name, mode, hard, soft, 1, 0 if soft > hard else 1, rng.choice(GEOMS + [(160, 2048), (2, 64)]), steps, rng,
                        mine_prob=rng.choice([1.0, 0.5, 0.3]), procs=procs, util=util, proc_hold=rng.choice([1, 1, 7, 31, 97]))


def gen_watcher_random():
    rng = random.Random(RANDOM_SEED)
    return [random_trajectory(rng, 100, "random%02d" % i) for i in range(RANDOM_TRAJECTORIES)]


ENV_CASES = [
    {"MANAGER_VISIBLE_DEVICES": helpers.STUB_UUID, "CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "10",
     "MANAGER_COMPATIBILITY_MODE": "0"},
    {"MANAGER_VISIBLE_DEVICES": helpers.STUB_UUID, "CUDA_MEM_LIMIT_0": "4g", "CUDA_CORE_LIMIT_0": "25"},
    {"MANAGER_VISIBLE_DEVICES": helpers.STUB_UUID, "CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4",
     "VMEMORY_NODE_ENABLED": "true", "MANAGER_COMPATIBILITY_MODE": "2"},
    {"MANAGER_VISIBLE_DEVICES": ",".join([helpers.STUB_UUID] + ["GPU-00000000-0000-0000-0000-000000000000"] * 3 +
                                         ["GPU-22222222-2222-2222-2222-222222222222"]),
     "CUDA_MEM_LIMIT": "1.5g", "CUDA_MEM_LIMIT_4": "512m", "CUDA_CORE_LIMIT": "30", "CUDA_CORE_SOFT_LIMIT_4": "80",
     "CUDA_MEM_OVERSOLD_4": "TRUE", "MANAGER_COMPATIBILITY_MODE": "101", "VGPU_POD_NAME": "pod-a",
     "VGPU_POD_NAMESPACE": "ns", "VGPU_POD_UID": "0123456789abcdef0123456789abcdef0123456789abcdefXYZ-too-long",
     "VGPU_CONTAINER_NAME": "main", "MANAGER_CLIENT_REGISTER_UUID": "reg-1", "EXTERNAL_SM_WATCHER_ENABLED": "1"},
    {"MANAGER_VISIBLE_DEVICE_0": helpers.STUB_UUID, "MANAGER_VISIBLE_DEVICE_1": "GPU-33333333-3333-3333-3333-333333333333",
     "CUDA_MEM_LIMIT": "2048k", "CUDA_CORE_LIMIT": "100", "CUDA_CORE_SOFT_LIMIT": "50"},
    {"NVIDIA_VISIBLE_DEVICES": helpers.STUB_UUID + ",GPU-44444444-4444-4444-4444-444444444444",
     "CUDA_MEM_LIMIT_1": "3t", "CUDA_MEM_RATIO": "0.5", "CUDA_MEM_OVERSOLD": "1", "CUDA_CORE_LIMIT_1": "0"},
    {"MANAGER_VISIBLE_DEVICES": helpers.STUB_UUID, "CUDA_MEM_LIMIT_0": "", "CUDA_MEM_LIMIT": "1g",
     "CUDA_CORE_LIMIT_0": "abc", "CUDA_MEM_RATIO_0": "1.000001", "VMEMORY_NODE_ENABLED": "yes"},
    {"MANAGER_VISIBLE_DEVICES": "GPU-aaaaaaaa-aaaa-aaaa-aaaa-aaaaaaaaaaaaTOOLONG-0123456789,," + helpers.STUB_UUID,
     "CUDA_MEM_LIMIT": "100", "CUDA_CORE_LIMIT": "7.9"},
    {},
]
ENV_KEYS = sorted({k for c in ENV_CASES for k in c} | {"NVIDIA_VISIBLE_DEVICES", "MANAGER_VISIBLE_DEVICES"})


def gen_env():
    out = []
    co = helpers.Cosim()
    for env in ENV_CASES:
        for k in ENV_KEYS:
            co.ask("unsetenv " + k)
        for k, v in env.items():
            assert co.ask("setenv %s %s" % (k, v)) == "ok"
        out.append({"env": env, "config_hex": co.ask("envcfg")})
    co.close()
    return out


def gen_used(rng):
    out = []
    # HOST mode: everything counts, graphics pids present in compute are dropped
    co = helpers.Cosim()
    for _ in range(40):
        nc, ng = rng.randint(0, 12), rng.randint(0, 8)
        comp = [[rng.randint(1, 30), rng.choice([0, 1, 1 << 20, rng.randint(0, 1 << 36)])] for _ in range(nc)]
        gfx = [[rng.randint(1, 30), rng.randint(0, 1 << 34)] for _ in range(ng)]
        line = "used 0 %d %s %d %s" % (nc, " ".join("%d:%d" % tuple(c) for c in comp), ng,
                                       " ".join("%d:%d" % tuple(g) for g in gfx))
        out.append({"mode": 0, "compute": comp, "graphics": gfx, "cflags": [], "gflags": [],
                    "used": int(co.ask(line))})
    # 2^64 wrap-around
    comp = [[1, (1 << 64) - 5], [2, 10]]
    out.append({"mode": 0, "compute": comp, "graphics": [], "cflags": [], "gflags": [],
                "used": int(co.ask("used 0 2 %d:%d %d:%d 0" % (1, comp[0][1], 2, 10)))})
    co.close()
    # cgroup v2 mode: membership from .host_proc/<pid>/cgroup
    for _ in range(25):
        sb = helpers.Sandbox()
        pids = list(range(100, 130))
        mine = {p: rng.random() < 0.5 for p in pids}
        for p in pids:
            if rng.random() < 0.1:
                continue  # no file at all -> not mine
            d = sb.path("etc/vgpu-manager/.host_proc/%d" % p)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "cgroup"), "w") as f:
                f.write("0::/\n" if mine[p] else "12:memory:/x\n0::/system.slice\n")
        for p in pids:
            if not os.path.exists(sb.path("etc/vgpu-manager/.host_proc/%d/cgroup" % p)):
                mine[p] = False
        co = helpers.Cosim(sb=sb)
        nc, ng = rng.randint(0, 10), rng.randint(0, 6)
        comp = [[rng.choice(pids), rng.randint(0, 1 << 33)] for _ in range(nc)]
        gfx = [[rng.choice(pids), rng.randint(0, 1 << 33)] for _ in range(ng)]
        line = "used 2 %d %s %d %s" % (nc, " ".join("%d:%d" % tuple(c) for c in comp), ng,
                                       " ".join("%d:%d" % tuple(g) for g in gfx))
        out.append({"mode": 2, "compute": comp, "graphics": gfx, "cflags": [1 if mine[c[0]] else 0 for c in comp],
                    "gflags": [1 if mine[g[0]] else 0 for g in gfx], "used": int(co.ask(line))})
        co.close()
    return out


def main():
    helpers.build_all()
    if not helpers.have_reference():
        sys.exit("oracle/_ref is missing: run this where /root/reference exists")
    rng = random.Random(SEED)
    co = helpers.Cosim()
    gold = {"seed": SEED, "delta": gen_delta(co, rng), "token": gen_token(co, rng), "rate": gen_rate(co, rng)}
    co.close()
    gold["env"] = gen_env()
    gold["used"] = gen_used(rng)
    with open(os.path.join(HERE, "limiter_memory.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    with open(os.path.join(HERE, "watcher.json"), "w") as f:
        json.dump({"seed": SEED, "trajectories": gen_watcher(rng)}, f, separators=(",", ":"))
    with open(os.path.join(HERE, "watcher_random.json"), "w") as f:
        json.dump({"seed": RANDOM_SEED, "trajectories": gen_watcher_random()}, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "limiter_memory.json"), os.path.join(HERE, "watcher.json"), os.path.join(HERE, "watcher_random.json"))


if __name__ == "__main__":
    main()
