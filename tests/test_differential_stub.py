"""CPU: the B200 library against the compiled reference library, same scripted tenant, same
stub driver, same (redirected) contract files.  Transcripts (return codes, reported numbers,
ledger contents) and the vgpu.config bytes each library publishes must be identical.

On the stub the "GPU" runs the library's kernels through the oracle (tests/stub/stubdrv.c), so
this suite checks the HOST logic - hook surface, request staging, locks, ledger file, error
conventions.  The kernels themselves are checked on a B200 in tests/test_gpu_parity.py.
"""
import os

import pytest

import helpers as H

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="oracle/_ref not built (no /root/reference)")

MiB = 1 << 20
GiB = 1 << 30
BASE = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID, "LOGGER_LEVEL": "0"}


def both(script, env, stub_env=None, args=(), prep=None):
    e = dict(BASE)
    e.update(env)
    e.update(stub_env or {})
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        if prep:
            prep(sb)
        out, err, _ = H.run_scenario(lib, script, e, sb=sb, args=args)
        cfg = sb.config_bytes()
        outs.append((out, cfg, err))
        sb.cleanup()
    return outs


def assert_same(outs):
    (a, ca, ea), (b, cb, eb) = outs
    assert a == b, "transcripts differ:\n--- reference\n%s\n--- b200\n%s\n%s" % (a, b, eb[-1500:])
    assert ca == cb, "published vgpu.config differs"
    assert len(ca) == 1848
    return a


def test_cap_1g_alloc_free_and_reported_numbers(built):
    script = "\n".join([
        "init 0", "totalmem", "meminfo", "nvmlinfo", "nvmlinfo2",
        "alloc %d" % (100 * MiB), "alloc %d" % (900 * MiB), "alloc 1", "alloc %d" % (24 * MiB), "alloc %d" % (24 * MiB - 1),
        "meminfo", "nvmlinfo", "nvmlinfo2", "free 0", "meminfo", "alloc %d" % (100 * MiB), "alloc %d" % (100 * MiB),
        "pitch 1000 1000 16", "pitch 1048576 2000 4", "allocasync %d" % MiB, "pool %d" % MiB, "pool %d" % GiB,
        "create %d" % (2 * MiB), "create %d" % GiB, "array 256 256 32 4", "array 4096 4096 32 4",
        "array3d 64 64 64 1 1", "mipmap 32 32 32 16 2", "meminfo", "nvmlinfo", "setmode 1", "persistence", "ledger 0",
    ]) + "\n"
    t = assert_same(both(script, {"CUDA_MEM_LIMIT_0": "1g"}))
    assert "alloc 1 -> 2" not in t.splitlines()[7]  # sanity: the script exercises both outcomes
    assert "-> 2 " in t  # at least one OOM happened


def test_oversold_spill_sequence_config4_shape(built):
    """BASELINE config 4 on the stub: 8 GiB virtual / 2 GiB physical, 128 x 64 MiB."""
    lines = ["init 0"]
    for i in range(130):
        lines.append("alloc %d" % (64 * MiB))
        if i in (0, 31, 32, 33, 95, 96, 127, 128):
            lines += ["meminfo", "nvmlinfo", "ledger 0"]
    lines += ["free 40", "free 5", "ledger 0", "meminfo", "managed %d 1" % MiB, "managed %d 2" % MiB, "ledger 0",
              "alloc %d" % (64 * MiB), "alloc %d" % (64 * MiB), "nvmlinfo2"]
    env = {"CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true"}
    t = assert_same(both("\n".join(lines) + "\n", env))
    rows = [l for l in t.splitlines() if l.startswith("alloc")]
    assert all("-> 0" in r for r in rows[:128]) and "-> 2" in rows[128]  # 32 GPU + 96 UVA, then OOM
    assert "ledger -> size 1 [self %d]" % (96 * 64 * MiB) in t


def test_oversold_without_ledger_counts_no_uva(built):
    """Default deployment (vmem_node off, Appendix B.14): spilled bytes are not counted."""
    lines = ["init 0"] + ["alloc %d" % (512 * MiB)] * 12 + ["meminfo", "nvmlinfo", "ledger 0"]
    assert_same(both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "2g", "CUDA_MEM_RATIO_0": "2"}))


def test_driver_oom_falls_back_to_uva_when_oversold(built):
    lines = ["init 0", "alloc %d" % (300 * MiB), "alloc %d" % (300 * MiB), "alloc %d" % (300 * MiB), "ledger 0", "meminfo",
             "pitch 4096 100000 4", "allocasync %d" % (300 * MiB), "ledger 0", "nvmlinfo"]
    env = {"CUDA_MEM_LIMIT_0": "4g", "CUDA_MEM_OVERSOLD_0": "true", "VMEMORY_NODE_ENABLED": "1"}
    t = assert_same(both("\n".join(lines) + "\n", env, {"STUB_PHYS_MEM": str(512 * MiB)}))
    assert "ledger -> size 1" in t


def test_other_tenants_and_graphics_dedup_host_mode(built):
    stub = {"STUB_OTHER_PROCS": "901:268435456:c,902:134217728:g,903:67108864:cg", "STUB_CTX_BYTES": str(300 * MiB)}
    lines = ["init 0", "meminfo", "nvmlinfo", "alloc %d" % (200 * MiB), "alloc %d" % (100 * MiB), "meminfo", "nvmlinfo", "nvmlinfo2"]
    assert_same(both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "1g"}, stub))


def test_cgroup_v2_membership(built):
    stub = {"STUB_OTHER_PROCS": "901:268435456:c,902:134217728:c,903:67108864:g"}

    def prep(sb):
        for pid, mine in ((901, True), (902, False), (903, True)):
            d = sb.path("etc/vgpu-manager/.host_proc/%d" % pid)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "cgroup"), "w") as f:
                f.write("0::/\n" if mine else "0::/kubepods/burstable/other\n")

    lines = ["init 0", "meminfo", "nvmlinfo", "alloc %d" % (600 * MiB), "alloc %d" % (200 * MiB), "meminfo"]
    env = {"CUDA_MEM_LIMIT_0": "1g", "MANAGER_COMPATIBILITY_MODE": "2"}
    t = assert_same(both("\n".join(lines) + "\n", env, stub, prep=prep))
    # own pid has no .host_proc entry => not counted; 901 + 903 are: 320 MiB used before any alloc
    assert "nvmlinfo -> 0 total 1073741824 free %d used %d" % (GiB - 320 * MiB, 320 * MiB) in t


def test_client_mode_pids_file(built):
    stub = {"STUB_OTHER_PROCS": "901:268435456:c,902:134217728:c"}

    def prep(sb):
        os.makedirs(sb.path("etc/vgpu-manager/registry"), exist_ok=True)
        client = sb.path("etc/vgpu-manager/registry/device-client")
        with open(client, "w") as f:
            f.write("#!/bin/sh\nexit 0\n")
        os.chmod(client, 0o755)
        with open(sb.path("etc/vgpu-manager/config/pids.config"), "w") as f:
            f.write("902\n17\n")

    lines = ["init 0", "meminfo", "alloc %d" % (900 * MiB), "alloc %d" % (800 * MiB), "nvmlinfo"]
    env = {"CUDA_MEM_LIMIT_0": "1g", "MANAGER_COMPATIBILITY_MODE": "200", "VGPU_POD_UID": "uid-1",
           "VGPU_CONTAINER_NAME": "c", "MANAGER_CLIENT_REGISTER_UUID": "r"}
    assert_same(both("\n".join(lines) + "\n", env, stub, prep=prep))


def test_no_limits_is_pass_through(built):
    lines = ["init 0", "totalmem", "meminfo", "nvmlinfo", "alloc %d" % GiB, "meminfo", "setmode 1", "persistence",
             "launch 100 2 2 2"]
    assert_same(both("\n".join(lines) + "\n", {}, {"STUB_TOTAL_MEM": str(16 * GiB)}))


def test_device_missing_from_config_is_fatal_at_cuInit(built):
    """init_device_cuda_cores (cuda_hook.c:500-503): a CUDA-visible device the config does not
    describe terminates the process at the first cuInit - same exit code, same message."""
    env = dict(BASE)
    env.update({"MANAGER_VISIBLE_DEVICES": "GPU-99999999-9999-9999-9999-999999999999", "CUDA_MEM_LIMIT_0": "1g",
                "CUDA_CORE_LIMIT_0": "10", "STUB_TOTAL_MEM": str(16 * GiB)})
    res = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        out, err, rc = H.run_scenario(lib, "init 0\ntotalmem\n", env, sb=sb, check=False)
        res.append((out, rc, "cuda device 0 cannot find the corresponding host device" in err))
        sb.cleanup()
    assert res[0] == res[1] == ("", 1, True)


def test_cuGetProcAddress_path_is_hooked(built):
    """cudart resolves everything through cuGetProcAddress; the cap must still apply."""
    lines = ["init 0", "meminfo", "alloc %d" % (2 * GiB), "alloc %d" % (512 * MiB), "meminfo", "totalmem"]
    t = assert_same(both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "1g"}, args=("--gpa",)))
    assert "alloc %d -> 2" % (2 * GiB) in t and "total %d" % GiB in t


def test_existing_config_file_wins_over_env(built):
    """A vgpu.config of exactly 1848 bytes dropped by the control plane is used as is."""
    cfg = H.Cfg()
    cfg.devices[0].uuid = H.STUB_UUID.encode()
    cfg.devices[0].total_memory = 3 * GiB
    cfg.devices[0].real_memory = 3 * GiB
    cfg.devices[0].memory_limit = 1
    cfg.devices[0].activate = 1
    raw = bytes(cfg)

    def prep(sb):
        with open(sb.path("etc/vgpu-manager/config/vgpu.config"), "wb") as f:
            f.write(raw)

    lines = ["init 0", "totalmem", "meminfo", "alloc %d" % (2 * GiB), "alloc %d" % (2 * GiB), "nvmlinfo"]
    outs = both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "1g"}, prep=prep)
    t = assert_same(outs)
    assert outs[0][1] == raw and "totalmem -> 0 %d" % (3 * GiB) in t


def test_nvml_only_client_gets_the_reference_numbers(built):
    """nvidia-smi style process: no CUDA context, so no place to run the quota kernel.  Like the reference
    (nvml_hook.c:47-103) the report is computed on the host there - same fold, same clamp - and no
    context is created behind the client's back.  Compared across compatibility modes, with other
    tenants' processes, graphics duplicates and a UVA ledger on the GPU."""
    script = "nvmlinit 0\nnvmlinfo\nnvmlinfo2\npersistence\nsetmode 0\n"
    cases = [
        {"CUDA_MEM_LIMIT_0": "1g", "STUB_OTHER_PROCS": "901:268435456:c"},
        {"CUDA_MEM_LIMIT_0": "1g", "STUB_OTHER_PROCS": "901:268435456:c,902:1048576:cg,903:4096:g,904:999999999999:c"},
        {"CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true",
         "STUB_OTHER_PROCS": "901:268435456:c,905:123456789:g"},
        {"CUDA_MEM_LIMIT_0": "2g", "MANAGER_COMPATIBILITY_MODE": "2", "STUB_OTHER_PROCS": "901:268435456:c,902:1048576:c"},
    ]
    for extra in cases:
        env = dict(BASE)
        env.update(extra)
        outs = []
        for lib in (H.REF_SO, H.NEW_SO):
            sb = H.Sandbox()
            if extra.get("MANAGER_COMPATIBILITY_MODE") == "2":  # pid 901 is a member of the container, 902 is not
                for pid, line in ((901, "0::/\n"), (902, "0::/kubepods/other\n")):
                    d = sb.path("etc/vgpu-manager/.host_proc/%d" % pid)
                    os.makedirs(d, exist_ok=True)
                    with open(os.path.join(d, "cgroup"), "w") as f:
                        f.write(line)
            out, err, _ = H.run_scenario(lib, script, dict(env, LOGGER_LEVEL="1"), sb=sb)
            sb.cleanup()
            outs.append(out)
        assert outs[0] == outs[1], (extra, outs)
        assert "nvmlinfo -> 0 " in outs[1]
    assert "used 268435456" in outs[1]


def test_scrub_on_free_runs_the_clear_kernel(built):
    """VGPU_B200_SCRUB_ON_FREE=1: the allocation is zeroed by vgpu_clear_kernel before cuMemFree."""
    script = "init 0\nalloc 1048576\ndirty 0 1048576\nfree 0\nalloc 4096\ndirty 1 4096\nfree 1\n"
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "STUB_CHECK_SCRUB": "1"})
    sb = H.Sandbox()
    _, err_off, _ = H.run_scenario(H.NEW_SO, script, env, sb=sb)
    sb.cleanup()
    sb = H.Sandbox()
    out, err_on, _ = H.run_scenario(H.NEW_SO, script, dict(env, VGPU_B200_SCRUB_ON_FREE="1"), sb=sb)
    sb.cleanup()
    assert err_off.count("freed dirty") == 2 and "freed clean" not in err_off
    assert err_on.count("freed clean") == 2 and "freed dirty" not in err_on
    assert "free h0 -> 0" in out and "free h1 -> 0" in out


def test_external_sm_watcher_file_is_mandatory_when_enabled(built):
    """EXTERNAL_SM_WATCHER_ENABLED without /etc/vgpu-manager/watcher/sm_util.config is fatal in
    both libraries (loader.c:2082-2087); with a well-formed 1 311 232-byte file both start."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "EXTERNAL_SM_WATCHER_ENABLED": "true"})
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        out, err, rc = H.run_scenario(lib, "init 0\ntotalmem\n", env, sb=sb, check=False)
        sb.cleanup()
        assert rc == 1 and out == "", (lib, rc, out)
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        with open(sb.path("etc/vgpu-manager/watcher/sm_util.config"), "wb") as f:
            f.write(b"\0" * 1311232)
        out, err, rc = H.run_scenario(lib, "init 0\ntotalmem\nmeminfo\n", env, sb=sb, check=False)
        outs.append((out, rc, sb.config_bytes()))
        sb.cleanup()
    assert outs[0] == outs[1] and outs[0][1] == 0


def _storm(lib, env, n, threads=1, extra=()):
    import json
    import subprocess
    sb = H.Sandbox()
    e = H.preload_env(lib, sb, env)
    r = subprocess.run([H.STORM, "--steps", "1", "--warmup", "0", "--per-step", str(n), "--threads", str(threads),
                        "--no-kernel"] + list(extra), env=e, capture_output=True, text=True, timeout=300)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_config1_stub_launch_storm_both_libraries(built):
    """BASELINE configs[0]: 10 % cores / 1 GiB, empty-kernel storm on the CPU-only stub driver,
    closed-loop utilisation model; single thread and 8 threads (the reference serialises every
    launch on one mutex, loader.c:1769).  Plumbing check: everything is admitted exactly once."""
    env = dict(BASE)
    env.update({"CUDA_CORE_LIMIT_0": "10", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "closed:0.02"})
    for threads in (1, 8):
        ref = _storm(H.REF_SO, env, 400000, threads)
        new = _storm(H.NEW_SO, env, 400000, threads)
        assert ref["launches"] == new["launches"] == 400000 and ref["fails"] == new["fails"] == 0
        assert new["limiter"]["present"] == 1 and new["sampler_launches"] > 0
        assert new["p50_ns"] < 5000


def test_every_launch_entry_point_is_forwarded_and_pays_the_reference_cost(built):
    """The 11 launch hooks (cuda_hook.c:1810-2002): kernel launches pay gridX*gridY*gridZ tokens, cuLaunchGrid[Async]
    width*height, the pre-CUDA-4 cuLaunch one token.  Same transcript as the reference through every entry point, and
    the B200 library's `consumed` counter moves by exactly that cost."""
    env = dict(BASE)
    env.update({"CUDA_CORE_LIMIT_0": "50", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "fixed:5"})
    n, gx, gy = 20, 7, 3
    script = "init 0\nlimstate\n" + "".join("launchvia %d %d %d %d\nlimstate\n" % (k, n, gx, gy) for k in range(9))
    ref, _, _ = H.run_scenario(H.REF_SO, script, env)
    new, _, _ = H.run_scenario(H.NEW_SO, script, env)
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("limstate")]
    assert strip(ref) == strip(new), (ref, new)
    assert sum("-> ok %d" % n in l for l in strip(new)) == 9, new
    consumed = [int(l.split()[2]) for l in new.splitlines() if l.startswith("limstate consumed")]
    assert len(consumed) == 10, new
    deltas = [b - a for a, b in zip(consumed, consumed[1:])]
    assert deltas == [n * gx * gy] * 8 + [n * 1], deltas


def test_hooked_blocking_calls_forward_under_a_core_cap(built):
    """cuCtxSynchronize and the synchronous copies are hooked so that a throttled thread waits in user
    space before it blocks inside the driver (limiter.c wait_until_unparked).  Plumbing check on the stub:
    a capped storm that issues one of them every 500 launches completes with the copy executed (the
    call's return code is part of `fails`), for both libraries - the reference forwards them."""
    env = dict(BASE)
    env.update({"CUDA_CORE_LIMIT_0": "10", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "closed:0.02"})
    for call in ("sync", "htod", "dtoh", "dtod", "copy"):
        for lib in (H.REF_SO, H.NEW_SO):
            d = _storm(lib, env, 100000, 2, ("--sync-every", "500", "--block-with", call))
            assert d["launches"] == 100000 and d["fails"] == 0, (call, lib, d)


def test_refill_does_not_need_the_host_to_enter_the_driver(built):
    """Regression for a deadlock class seen with PyTorch on the real driver: a thread blocks inside
    a driver call (holding the context lock) behind a parked stream, so no other thread can launch
    the kernel that would refill the bucket.  STUB_CTX_LOCK=1 models that lock in the fake driver;
    the parked stream must be released without a host launch - by the watchdog's loan in the
    default mode, by the resident governor in VGPU_B200_GOVERNOR=1 mode."""
    for governor in ("0", "1"):
        env = dict(BASE)
        env.update({"CUDA_CORE_LIMIT_0": "10", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "closed:0.02", "STUB_CTX_LOCK": "1",
                    "VGPU_B200_GOVERNOR": governor})
        for threads in (1, 4):
            new = _storm(H.NEW_SO, env, 300000, threads)
            assert new["launches"] == 300000 and new["fails"] == 0
            assert new["limiter"]["present"] == 1 and new["gated_launches"] > 0, new
            if governor == "0":  # the tick thread is locked out whenever a stream is parked: only loans release it
                assert new["watchdog_loans"] > 0, new


def test_directly_linked_tenant_is_intercepted_by_symbol_interposition(built):
    """No dlopen/dlsym/cuGetProcAddress at all: the tenant links libcuda; the preloaded library
    must win plain symbol resolution for every hooked entry point."""
    import subprocess
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "50", "STUB_UTIL": "closed:0.02"})
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        r = subprocess.run([os.path.join(H.BUILD, "direct")], env=H.preload_env(lib, sb, env), capture_output=True,
                           text=True, timeout=120)
        sb.cleanup()
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(r.stdout)
    assert outs[0] == outs[1]
    assert "totalmem 0 1073741824" in outs[0] and "alloc 0\nalloc 2\n" in outs[0] and "launch 1000" in outs[0]


def test_forked_child_is_served_like_the_parent(built):
    """fork awareness (loader.c:1805-1822,2054-2066): config and index maps are reloaded in the
    child.  Memory path compared against the reference; launches in the child only under the
    B200 library - the reference never restarts its watcher there and would hang a capped child
    (SURVEY.md Appendix B.13), the B200 library re-arms its tick thread."""
    env = {"CUDA_MEM_LIMIT_0": "1g"}
    script = "init 0\nalloc %d\nforkchild %d 0\nmeminfo\nforkchild %d 0\n" % (300 * MiB, 200 * MiB, 900 * MiB)
    t = assert_same(both(script, env))
    assert "child alloc 0 meminfo 0 total 1073741824" in t and "child alloc 2 " in t
    e = dict(BASE)
    e.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "20", "STUB_UTIL": "closed:0.02"})
    sb = H.Sandbox()
    out, err, _ = H.run_scenario(H.NEW_SO, "init 0\nlaunch 5000 1 1 1\nforkchild 1048576 200000\nlaunch 5000 1 1 1\n", e, sb=sb)
    sb.cleanup()
    assert "child alloc 0 meminfo 0 total 1073741824 launches 200000" in out and out.count("-> ok 5000") == 2


def test_cuda_graph_replays_forward_by_default_and_are_metered_on_request(built):
    """cuGraphLaunch is a pure forward in the reference (cuda_originals.c:3033) and by default here:
    same transcripts.  With VGPU_B200_GRAPH_LIMIT=1 a replay pays the sum of its kernel nodes' grids
    (dlsym and cuGetProcAddress bindings alike) and is gated like a launch."""
    script = "init 0\ngraph 10 100\ngraphlaunch 50\ngraph 3 7\ngraphlaunch 4\nlaunch 5 7 1 1\n"
    env = {"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "50"}
    for args in ((), ("--gpa",)):
        assert_same(both(script, env, args=args))
    e = dict(BASE)
    e.update(env)
    metered = dict(e, VGPU_B200_GRAPH_LIMIT="1")
    for args in ((), ("--gpa",)):
        out, _, sb = H.run_scenario(H.NEW_SO, script + "limstate\n", e, args=args)
        sb.cleanup()
        assert "limstate consumed 35" in out  # only the five plain launches
        out, _, sb = H.run_scenario(H.NEW_SO, script + "limstate\n", metered, args=args)
        sb.cleanup()
        assert "graphlaunch 50 -> ok 50" in out and "graphlaunch 4 -> ok 4" in out
        assert "limstate consumed %d" % (50 * 10 * 100 + 4 * 3 * 7 + 35) in out
    # saturating tenant under a 10 % cap: replays end up behind the gate
    hot = dict(metered, CUDA_CORE_LIMIT_0="10", STUB_UTIL="fixed:90", LOGGER_LEVEL="3")
    out, err, sb = H.run_scenario(H.NEW_SO, "init 0\ngraph 8 4000\ngraphlaunch 400\n", hot)
    sb.cleanup()
    assert "graphlaunch 400 -> ok 400" in out and "metric=rate_gated" in err


def test_device_reset_rebuilds_the_device_state(built):
    """cudaDeviceReset() mid-process: the reference keeps no device state and simply carries on;
    this library's token bucket / slab / streams die with the context (the stub unmaps them, so a
    stale access faults) and must be rebuilt on next use - same transcript, cap still enforced,
    launches still metered."""
    script = ("init 0\nalloc %d\nalloc %d\nmeminfo\nlaunch 300 4 1 1\n"
              "reset\nmeminfo\nalloc %d\nalloc %d\nalloc %d\nmeminfo\nnvmlinfo\nlaunch 300 4 1 1\n"
              "reset\nalloc %d\nmeminfo\n") % (256 * MiB, 256 * MiB, 512 * MiB, 400 * MiB, 400 * MiB, 900 * MiB)
    env = {"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "50", "STUB_UTIL": "fixed:20"}
    for args in ((), ("--gpa",)):
        t = assert_same(both(script, env, args=args))
        assert t.count("reset 0 0 0") == 2 and "-> 2" in t  # the cap is enforced again after the reset
    e = dict(BASE)
    e.update(env)
    out, err, sb = H.run_scenario(H.NEW_SO, script + "limstate\n", dict(e, LOGGER_LEVEL="3"))
    sb.cleanup()
    assert err.count("device runtime up") == 3 and err.count("will be rebuilt on next use") == 2, err[-1500:]
    assert "limstate consumed 0" in out  # a fresh bucket after the last reset (nothing launched since)


def test_uva_records_outlive_a_device_reset_like_the_reference_list(built):
    """The reference's list of UVA allocations is host memory: a device reset leaves its nodes in place, and when a
    later allocation that gets the same address is freed, the stale node is found and the ledger shrinks by ITS size
    (loader.c:1869-1907) - also when the new allocation was not a UVA one.  The records here live in HBM; they are
    read back while the context is torn down and searched after the live table (found by the offline sweep: a 1-byte
    ledger difference in 2 of 2400 random scripts).  What cannot be matched in general is WHICH later allocation gets a
    stale address: the library's own device allocations perturb the driver's address sequence, so in long random
    scripts the coincidence sometimes happens under one library only (3 of 3400 scripts in the later sweeps); here
    both lives bring the runtime up at the same point, so the addresses repeat under both."""
    env = {"CUDA_MEM_LIMIT_0": "4g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true", "STUB_PHYS_MEM": str(8 * GiB + 48 * MiB)}
    # a VMM handle fills the 1 GiB physical share, so the two small allocations are UVA ones (ledger 12288); after the
    # reset the same two requests get the same addresses on the GPU path (ledger untouched), and freeing them removes
    # the stale nodes: 12288 -> 8192 -> 0.  Then a second life with live records over stale addresses (newest first).
    script = ("init 0\ncreate %d\nalloc 4096\nalloc 8192\nledger 0\nreset\nledger 0\n"
              "alloc 4096\nalloc 8192\nledger 0\nfree 0\nledger 0\nfree 1\nledger 0\n"
              "create %d\nalloc 4096\nledger 0\nreset\ncreate %d\nalloc 4096\nledger 0\nfree 1\nledger 0\n"
              "reset\nalloc 4096\nfree 0\nledger 0\nnvmlinfo\n") % (GiB, GiB, GiB)
    for args in ((), ("--gpa",)):
        outs = both(script, env, args=args)
        ref_led = [int(l.split()[-1].rstrip("]")) for l in outs[0][0].splitlines() if l.startswith("ledger")]
        if ref_led[:5] != [12288, 12288, 12288, 8192, 0]:
            pytest.skip("the fake driver (mmap) did not hand the same addresses out again after the reset on this kernel")
        t = assert_same(outs)
        led = [int(l.split()[-1].rstrip("]")) for l in t.splitlines() if l.startswith("ledger")]
        assert led[:5] == [12288, 12288, 12288, 8192, 0], led   # stale nodes matched by address after the reset
        assert led[5:] == [4096, 8192, 4096, 0], led            # live record first, then the stale one under the same address


def test_one_process_on_two_gpus_with_a_hole_in_the_visible_list(built):
    """MANAGER_VISIBLE_DEVICES lists host GPUs by index with all-zero UUIDs as holes (util.c:27-120):
    host 0 = CUDA device 1, host 1 = hole, host 2 = CUDA device 0.  One process uses both devices;
    caps, reported numbers and the written vgpu.config must follow the host index, not the CUDA one."""
    u0, u1 = H.STUB_UUID, "GPU-22222222-2222-2222-2222-222222222222"
    hole = "GPU-00000000-0000-0000-0000-000000000000"
    env = {"MANAGER_VISIBLE_DEVICES": ",".join((u1, hole, u0)), "CUDA_MEM_LIMIT_0": "1g", "CUDA_MEM_LIMIT_2": "3g",
           "CUDA_CORE_LIMIT_0": "30", "CUDA_CORE_LIMIT_2": "60", "CUDA_MEM_LIMIT": "2g"}
    script = ("init 0\ntotalmem\nmeminfo\nalloc %d\nalloc %d\nmeminfo\nnvmlinfo\n"        # CUDA 0 = host 2: 3 GiB cap
              "dev 1\ntotalmem\nmeminfo\nalloc %d\nalloc %d\nmeminfo\nnvmlinfo\nlaunch 200 3 1 1\n"  # CUDA 1 = host 0: 1 GiB cap
              "dev 0\nmeminfo\nalloc %d\nmeminfo\nlaunch 200 3 1 1\n") % (GiB, GiB, 600 * MiB, 600 * MiB, 2 * GiB)
    outs = both(script, env, {"STUB_GPU_COUNT": "2", "STUB_UTIL": "fixed:10"})
    t = assert_same(outs)
    lines = t.splitlines()
    assert lines[1].endswith(str(3 * GiB)) and any(l.startswith("totalmem") and l.endswith(str(GiB)) for l in lines[8:11]), t
    assert t.count("-> 2") == 2  # one refusal per device, each against its own cap


def test_two_gpus_frees_and_reports_across_devices(built):
    """Three behaviours an offline two-GPU fuzz sweep (tests/fuzz_sweep.py) found differing, pinned here:
    (1) the reference keeps ONE list of UVA allocations per process and, on free, subtracts the node from the ledger
        record of whatever device is *current* (loader.c:1869-1907) - an allocation ledgered under GPU A and freed while
        GPU B is current leaves A's record untouched and shrinks B's;
    (2) in slab mode a slab may likewise be freed while another device is current;
    (3) a device without a memory cap reports the driver's own numbers - the library's device-resident state (brought up
        there by a core limit or by a ledgered managed allocation) must not show in them."""
    u0, u1 = H.STUB_UUID, "GPU-22222222-2222-2222-2222-222222222222"
    stub = {"STUB_GPU_COUNT": "2", "STUB_UTIL": "fixed:5"}
    # (1) both GPUs oversold with a ledger; h1 spills on GPU 0, is freed on GPU 1, which has a spilled allocation of its own
    env = {"MANAGER_VISIBLE_DEVICES": u0 + "," + u1, "CUDA_MEM_LIMIT_0": "1g", "CUDA_MEM_RATIO_0": "4", "CUDA_MEM_LIMIT_1": "1g",
           "CUDA_MEM_RATIO_1": "4", "VMEMORY_NODE_ENABLED": "true"}
    script = ("init 0\nalloc %d\nalloc %d\nledger 0\ndev 1\nalloc %d\nalloc %d\nledger 1\nfree 1\nledger 0\nledger 1\nnvmlinfo\n"
              "dev 0\nnvmlinfo\nfree 3\nledger 0\nledger 1\nmeminfo\n") % (200 * MiB, 128 * MiB, 200 * MiB, 96 * MiB)
    t = assert_same(both(script, env, stub))
    led = [l for l in t.splitlines() if l.startswith("ledger")]
    assert led[0].endswith("[self %d]" % (128 * MiB)) and led[1].endswith("[self %d]" % (96 * MiB)), t
    assert led[2] == led[0] and led[3].endswith("[self 0]"), t          # GPU 0 keeps its record, GPU 1 paid for the free
    # (2) the same hop in slab mode (ignored by the reference)
    t2 = assert_same(both(script, dict(env, VGPU_B200_SLAB="1"), stub))
    assert t2 == t
    # (3) GPU 1 has only a core limit; GPU 0 nothing at all but sees a GLOBAL-attached managed allocation
    env3 = {"MANAGER_VISIBLE_DEVICES": u0 + "," + u1, "CUDA_CORE_LIMIT_1": "10", "VMEMORY_NODE_ENABLED": "true"}
    script3 = ("init 0\nmeminfo\nmanaged 4096 1\nmeminfo\nnvmlinfo\nnvmlinfo2\ndev 1\nmeminfo\nlaunch 20 1 1 1\nalloc %d\nmeminfo\nnvmlinfo\n"
               "nvmlinfo2\n") % (64 * MiB)
    assert_same(both(script3, env3, stub))


def test_reset_of_one_gpu_leaves_the_other_gpus_runtime_alone(built):
    """Two GPUs, both capped; the tenant resets the primary context of one of them.  Only that device's runtime goes away
    (and its share of the footprint registry with it); allocations, reports and the ledger on the other device carry on,
    and the reset device is rebuilt at its next use - transcript identical to the reference."""
    u0, u1 = H.STUB_UUID, "GPU-22222222-2222-2222-2222-222222222222"
    env = {"MANAGER_VISIBLE_DEVICES": u1 + "," + u0, "CUDA_MEM_LIMIT_0": "1g", "CUDA_MEM_RATIO_0": "2", "CUDA_MEM_LIMIT_1": "1g",
           "VMEMORY_NODE_ENABLED": "true"}
    script = ("init 0\nmanaged %d 1\ndev 1\nalloc %d\nmeminfo\ndev 0\nreset\ndev 1\nalloc 4096\nmeminfo\nnvmlinfo\nledger 0\n"
              "dev 0\nmeminfo\nalloc %d\nmeminfo\nnvmlinfo\nledger 1\n") % (200 * MiB, 64 * MiB, 32 * MiB)
    outs = both(script, env, {"STUB_GPU_COUNT": "2", "STUB_UTIL": "fixed:5"})
    assert_same(outs)
    # an un-capped GPU that only carried a ledgered managed allocation: after its reset nothing of the library is left there
    env2 = {"MANAGER_VISIBLE_DEVICES": u0 + "," + u1, "CUDA_MEM_LIMIT_1": "1g", "VMEMORY_NODE_ENABLED": "true"}
    script2 = "init 0\nmanaged 4096 1\nnvmlinfo\nreset\ndev 0\nnvmlinfo\nmeminfo\ndev 1\nalloc 4096\nnvmlinfo\n"
    assert_same(both(script2, env2, {"STUB_GPU_COUNT": "2", "STUB_UTIL": "fixed:5"}))


def test_closed_loop_share_is_in_the_references_ballpark(built):
    """The reference defines no core-% tolerance (SURVEY.md 8a L-tol); what can be compared is the rate each
    limiter settles at when the fake GPU's utilisation follows the tenant's own launch rate (closed loop,
    same model for both).  Measured: 0.84-1.31 x the reference over caps 10/25/50 %; asserted loosely."""
    import json
    import subprocess
    rates = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        e = dict(BASE)
        e.update({"CUDA_CORE_LIMIT_0": "10", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "closed:0.02"})
        r = subprocess.run([H.STORM, "--steps", "1000", "--warmup", "0", "--per-step", "20000", "--no-kernel", "--max-seconds", "6"],
                           env=H.preload_env(lib, sb, e), capture_output=True, text=True, timeout=120)
        sb.cleanup()
        assert r.returncode == 0, r.stderr[-1500:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        rates.append(d["launches"] / d["wall_s"])
    assert 0.5 < rates[1] / rates[0] < 2.0, rates


def test_eight_threads_allocating_under_an_oversold_cap(built):
    """Thread safety of the memory path: 8 threads x 400 steps of alloc / managed alloc / free / report / launch in
    one process (384 MiB virtual, 192 MiB physical, ledger on).  Per-call outcomes depend on the interleaving;
    what must hold under either library: only SUCCESS or OUT_OF_MEMORY, self-consistent reports, and - once
    every thread has freed what it held - the same final figures."""
    import subprocess
    finals = []
    for lib in (H.REF_SO, H.NEW_SO):
        for seed in (1, 2):
            sb = H.Sandbox()
            e = dict(BASE)
            e.update({"CUDA_MEM_LIMIT_0": "384m", "CUDA_MEM_RATIO_0": "2", "VMEMORY_NODE_ENABLED": "true", "CUDA_CORE_LIMIT_0": "50",
                      "STUB_UTIL": "fixed:10"})
            r = subprocess.run([os.path.join(H.BUILD, "mtstorm"), "--threads", "8", "--steps", "400", "--seed", str(seed)],
                               env=H.preload_env(lib, sb, e), capture_output=True, text=True, timeout=120)
            sb.cleanup()
            assert r.returncode == 0, (lib, r.stdout, r.stderr[-1500:])
            assert "oom 0\n" not in r.stderr  # the cap was actually hit
            finals.append(r.stdout)
    assert len(set(finals)) == 1, finals


def test_ledger_hygiene_purges_dead_processes_identically(built):
    """vmem_node.config outlives its writers: at start-up the library drops records of dead or zombie pids
    (swap-with-last removal, loader.c:1580-1673) and at exit its own.  Starting from the same dirty file, both
    libraries must leave byte-identical files behind."""
    import struct
    dead = [4100001, 4100002, 4100003, 4100004]
    alive = [1, os.getpid()]

    def dirty(sb):
        raw = bytearray(262272)
        def put(dev, recs):
            base = dev * 16392
            for i, (pid, used) in enumerate(recs):
                struct.pack_into("<iiQ", raw, base + 16 * i, pid, 0, used)
            struct.pack_into("<I", raw, base + 16384, len(recs))
        put(0, [(dead[0], 5 * MiB), (alive[0], 7 * MiB), (dead[1], 11 * MiB), (alive[1], 13 * MiB), (dead[2], 17 * MiB)])
        put(3, [(dead[3], 19 * MiB), (alive[1], 23 * MiB)])
        put(5, [(alive[0], 29 * MiB)])
        with open(sb.ledger(), "wb") as f:
            f.write(raw)

    script = "init 0\nledger 0\n" + "alloc %d\n" % (300 * MiB) * 3 + "ledger 0\nmeminfo\nnvmlinfo\n"
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_MEM_RATIO_0": "2", "VMEMORY_NODE_ENABLED": "true"})
    files, outs = [], []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        dirty(sb)
        out, err, _ = H.run_scenario(lib, script, env, sb=sb)
        files.append(open(sb.ledger(), "rb").read())
        outs.append(out)
        sb.cleanup()
    assert outs[0] == outs[1]
    assert files[0] == files[1]
    n0 = struct.unpack_from("<I", files[1], 16384)[0]
    pids0 = sorted(struct.unpack_from("<iiQ", files[1], 16 * i)[0] for i in range(n0))
    assert pids0 == sorted(alive)  # the dead are gone, the tenant removed itself at exit


def test_cuDriverGetVersion_alone_publishes_the_config(built):
    """Appendix B.12: the cuDriverGetVersion hook loads the config and the device mapping (so vgpu.config is
    written from env) without cuInit and without starting the watcher."""
    for args in ((), ("--gpa",)):
        outs = both("drvver\n", {"CUDA_MEM_LIMIT_0": "1.5g", "CUDA_CORE_LIMIT_0": "30", "CUDA_CORE_SOFT_LIMIT_0": "60"}, args=args)
        t = assert_same(outs)
        assert t == "drvver -> 0 12090\n"
        cfg = H.Cfg.from_buffer_copy(outs[1][1])
        assert cfg.devices[0].total_memory == 1536 * MiB and cfg.devices[0].hard_core == 30 and cfg.devices[0].soft_core == 60


def test_idle_tenant_keeps_no_kernel_on_the_gpu(built):
    """(On-device queue signal, VGPU_B200_UTIL_SOURCE=queue.)
    While nothing of the tenant is executing the tick thread launches no sampler windows at all (an idle
    or fully throttled tenant must not hold a time slice of a shared GPU); the ticks it skipped are
    accounted as idle windows and the control periods among them are replayed by the next launch."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:20", "VGPU_B200_SKIP_IDLE_WINDOWS": "1",
                "VGPU_B200_UTIL_SOURCE": "queue"})
    script = "init 0\nlaunch 200 4 1 1\nsleepms 1200\nmetrics 0\n"
    out, err, sb = H.run_scenario(H.NEW_SO, script, env)
    sb.cleanup()
    m = [l for l in out.splitlines() if l.startswith("metrics")][0].split()
    launches, skipped, steps = int(m[2]), int(m[4]), int(m[6])
    # ~120 ticks went by; on the fake GPU tenant kernels finish instantly, so nearly all of them are skipped
    assert skipped >= 80 and launches <= 5, out
    # same script with skipping off: one launch per tick and a control step every 8th
    out2, _, sb = H.run_scenario(H.NEW_SO, script, dict(env, VGPU_B200_SKIP_IDLE_WINDOWS="0"))
    sb.cleanup()
    m2 = [l for l in out2.splitlines() if l.startswith("metrics")][0].split()
    assert int(m2[2]) >= 80 and int(m2[4]) == 0 and int(m2[6]) >= 10, out2


def test_default_reading_is_one_refill_launch_per_control_period(built):
    """Default utilisation source (the reference's NVML samples folded on the device): exactly one
    library kernel - vgpu_refill_kernel - per 80 ms control period, idle tenant or not."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:20"})
    out, err, sb = H.run_scenario(H.NEW_SO, "init 0\nlaunch 200 4 1 1\nsleepms 1200\nmetrics 0\n", env)
    sb.cleanup()
    m = [l for l in out.splitlines() if l.startswith("metrics")][0].split()
    launches, skipped, steps = int(m[2]), int(m[4]), int(m[6])
    assert 10 <= launches <= 17 and skipped == 0 and steps == launches, out


SLAB_SCRIPT = """init 0
nvmlinit 0
alloc 67108864
fill 0 67108864 17
alloc 67108864
fill 1 67108864 34
alloc 67108864
fill 2 67108864 51
alloc 67108864
fill 3 67108864 68
nvmlinfo
ledger 0
alloc 67108864
fill 4 67108864 85
nvmlinfo
meminfo
ledger 0
alloc 67108864
fill 5 67108864 102
alloc 33554432
fill 6 33554432 119
nvmlinfo
ledger 0
check 0 67108864 17
check 1 67108864 34
check 2 67108864 51
check 3 67108864 68
check 4 67108864 85
check 5 67108864 102
check 6 33554432 119
free 4
nvmlinfo
ledger 0
check 0 67108864 17
check 5 67108864 102
free 1
nvmlinfo
ledger 0
check 0 67108864 17
check 2 67108864 51
check 3 67108864 68
check 5 67108864 102
free 0
free 5
free 6
nvmlinfo
ledger 0
check 2 67108864 51
check 3 67108864 68
alloc 1073741824
free 2
free 3
nvmlinfo
ledger 0
"""


def test_slab_mode_spills_with_reference_accounting(built):
    """VGPU_B200_SLAB=1 on an oversold device (1 GiB cap over 256 MiB physical, ledger on): every
    reported number - NVML view, cuMemGetInfo, ledger bytes, OOM point - is the reference's, while the
    data really moves: the 5th and 6th 64 MiB allocation each demote the coldest HBM slab to host
    memory (spill copy) and take its scrubbed HBM; a 32 MiB allocation finds no victim of its size class
    and is host-backed itself; frees bring the partner slab home (promote) or push it out (demote).
    Every buffer keeps its contents through all of it."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true", "LOGGER_LEVEL": "1"})
    ref, _, sb = H.run_scenario(H.REF_SO, SLAB_SCRIPT, env)
    sb.cleanup()
    plain, _, sb = H.run_scenario(H.NEW_SO, SLAB_SCRIPT, env)
    sb.cleanup()
    slab, err, sb = H.run_scenario(H.NEW_SO, SLAB_SCRIPT + "slabstats 0\n", dict(env, VGPU_B200_SLAB="1"))
    sb.cleanup()
    assert plain == ref
    lines = slab.splitlines()
    assert "\n".join(lines[:-1]) + "\n" == ref, err[-2000:]
    assert "CORRUPT" not in slab and slab.count("intact") == 15
    st = dict(zip(lines[-1].split()[1::2], map(int, lines[-1].split()[2::2])))
    MiB64 = 64 << 20
    # allocs: 7 slabs (the 1 GiB request is refused by the cap before any slab exists)
    assert st["allocs"] == 7 and "alloc 1073741824 -> 2" in slab
    # demotions: h4 and h5 each push one slab out; freeing h1 (a demoted, GPU-accounted slab) pushes h5's HBM out too
    assert st["demotions"] == 3 and st["spill_bytes"] == 3 * MiB64
    # freeing h4 (UVA-accounted, in HBM) brings its demoted partner home
    assert st["promote_bytes"] == MiB64
    # HBM taken over from a victim is scrubbed before the newcomer sees it
    assert st["scrubbed_bytes"] == 2 * MiB64


def test_enforcement_tunables_are_ignored_under_a_mounted_config(built):
    """The tenant controls its environment.  When the control plane has mounted a vgpu.config the
    enforcement-affecting VGPU_B200_* knobs must not be honoured (they are only when the limits came from
    the tenant's own env anyway).  First run: env-built config, knob honoured (queue signal: a sampler
    window per 10 ms tick).  Second run in the same sandbox: the config file now exists = "mounted" ->
    default reading (one refill launch per 80 ms control period) although the knob is still set."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:20",
                "VGPU_B200_UTIL_SOURCE": "queue", "VGPU_B200_SKIP_IDLE_WINDOWS": "0"})
    script = "init 0\nlaunch 50 4 1 1\nsleepms 1000\nmetrics 0\n"
    sb = H.Sandbox()
    out1, _, _ = H.run_scenario(H.NEW_SO, script, env, sb=sb)
    assert os.path.exists(sb.path("etc/vgpu-manager/config/vgpu.config"))
    out2, _, _ = H.run_scenario(H.NEW_SO, script, env, sb=sb)
    sb.cleanup()

    def counts(out):
        m = [l for l in out.splitlines() if l.startswith("metrics")][0].split()
        return int(m[2]), int(m[6])
    l1, s1 = counts(out1)
    l2, s2 = counts(out2)
    assert l1 >= 60 and l1 >= 5 * s1, (l1, s1)      # queue signal: ~100 windows, a step every 8th
    assert 8 <= l2 <= 16 and l2 == s2, (l2, s2)     # mounted config: the knob is ignored


def test_control_plane_sets_tunables_through_a_file_next_to_the_mounted_config(built):
    """Under a mounted vgpu.config the knobs belong to the control plane: an optional b200.tunables in the same
    (read-only) config directory, NAME=value per line.  The tenant's environment says `nvml`, the file says `queue`:
    the file wins; comments, unknown names and a name that is only a prefix of a line are ignored."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:20", "VGPU_B200_SKIP_IDLE_WINDOWS": "0"})
    script = "init 0\nlaunch 50 4 1 1\nsleepms 1000\nmetrics 0\n"
    sb = H.Sandbox()
    H.run_scenario(H.NEW_SO, "init 0\n", env, sb=sb)  # leaves the env-built vgpu.config behind = "mounted" from now on
    assert os.path.exists(sb.path("etc/vgpu-manager/config/vgpu.config"))
    with open(sb.path("etc/vgpu-manager/config/b200.tunables"), "w") as f:
        f.write("# node defaults\nVGPU_B200_UTIL_SOURCE_X=sm\nSOMETHING=else\nVGPU_B200_UTIL_SOURCE=queue\nVGPU_B200_SKIP_IDLE_WINDOWS=0\n")
    env["VGPU_B200_UTIL_SOURCE"] = "nvml"
    out, _, _ = H.run_scenario(H.NEW_SO, script, env, sb=sb)
    assert os.path.exists(sb.path("lock/vgpu_0.readings"))  # and a tenant on an on-device signal publishes its reading (8f-1)
    sb.cleanup()
    m = [l for l in out.splitlines() if l.startswith("metrics")][0].split()
    launches, steps = int(m[2]), int(m[6])
    assert launches >= 60 and launches >= 5 * steps, (launches, steps)  # the queue signal's cadence, not the refill's


def test_failed_bring_up_is_retried_and_can_fail_closed(built):
    """A transient bring-up failure (module load on a full GPU): partial state is freed, capped allocations are refused
    (NOT_SUPPORTED) and the next hooked call after the 1 s back-off brings the runtime up.  There is no CPU enforcement
    path: by default launches pass un-throttled meanwhile (ERROR log); with VGPU_B200_FAIL_CLOSED=1 they are refused too."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:5", "STUB_FAIL_MODULE": "1", "LOGGER_LEVEL": "2"})
    script = "init 0\nlaunch 5 1 1 1\nalloc 4096\nsleepms 1200\nlaunch 5 1 1 1\nalloc 4096\nmeminfo\nmetrics 0\n"
    for closed, first in ((None, "launch 5 -> ok 5"), ("1", "launch 5 -> ok 0")):
        e = dict(env)
        if closed:
            e["VGPU_B200_FAIL_CLOSED"] = closed
        out, err, sb = H.run_scenario(H.NEW_SO, script, e)
        sb.cleanup()
        lines = out.splitlines()
        assert lines[1] == first and lines[2] == "alloc 4096 -> 801 h-1", out
        assert lines[4] == "launch 5 -> ok 5" and lines[5] == "alloc 4096 -> 0 h0", out  # after the back-off: up, metered, capped
        assert "bring-up failed on cuda device 0 (attempt 1)" in err and err.count("bring-up failed") == 1


def test_watcher_runs_from_cuinit_and_the_controller_catches_up(built):
    """The reference's watcher steps from the first successful cuInit (cuda_hook.c:566-577), i.e. while the
    application is still creating its context.  The device-resident controller can only exist once a
    context does, so the tick thread keeps the publications of the periods that elapse before that and
    replays them through vgpu_refill_kernel the moment the runtime is up: one second between cuInit and
    the first launch => about twelve control steps already taken when the first launch returns."""
    env = dict(BASE)
    env.update({"CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:3", "LOGGER_LEVEL": "3"})
    out, err, sb = H.run_scenario(H.NEW_SO, "init 0\nsleepms 1000\nlaunch 3 1 1 1\nsleepms 30\nmetrics 0\n", env)
    sb.cleanup()
    m = [l for l in out.splitlines() if l.startswith("metrics")][0].split()
    launches, steps = int(m[2]), int(m[6])
    assert 10 <= steps <= 16 and launches == steps, out
    assert "replayed" in err and "control periods that elapsed before the device runtime came up" in err
