"""GPU (B200): the B200 library against the compiled reference library on the REAL driver.

Same tenant script, one process per library, same cap; return codes and every reported number
(cuMemGetInfo, cuDeviceTotalMem, nvmlDeviceGetMemoryInfo[_v2], ledger) must match.  On real
hardware `used` is NVML's per-process figure, which includes the CUDA context's own footprint;
the B200 library additionally owns a small HBM block + module, which its quota kernel removes
again (self_bytes) - this test is what proves that compensation exact.
"""
import json
import os
import subprocess

import pytest

import helpers as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref/libvgpu-control.so did not travel")]
MiB = 1 << 20
GiB = 1 << 30


def gpu0_uuid():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return out.stdout.splitlines()[0].strip()


def both(script, env, args=()):
    base = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "LOGGER_LEVEL": "1",
            "CUDA_VISIBLE_DEVICES": "0"}
    base.update(env)
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        out, err, _ = H.run_scenario(lib, script, base, sb=sb, stub=False, args=args, timeout=300)
        outs.append((out, err))
        sb.cleanup()
    return outs


def test_memory_cap_numbers_match_reference_on_real_driver(built):
    lines = ["init 0", "totalmem", "meminfo", "nvmlinfo", "nvmlinfo2", "alloc %d" % (256 * MiB), "meminfo", "nvmlinfo",
             "alloc %d" % (1 * GiB), "alloc %d" % (2 * GiB), "alloc %d" % (1 * GiB), "meminfo", "nvmlinfo", "nvmlinfo2",
             "free 0", "meminfo", "pitch 4096 4096 16", "create %d" % (64 * MiB), "array 1024 1024 32 4", "meminfo",
             "nvmlinfo", "setmode 1", "persistence"]
    (a, ea), (b, eb) = both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "4g"})
    assert a == b, "reference:\n%s\nb200:\n%s\n%s" % (a, b, eb[-2000:])
    assert "-> 2" in a  # the 4 GiB cap was hit somewhere


def test_oversold_spill_sequence_matches_reference_on_real_driver(built):
    """BASELINE config 4: 8 GiB virtual / 2 GiB physical, 64 MiB allocations until OOM."""
    lines = ["init 0"]
    for i in range(132):
        lines.append("alloc %d" % (64 * MiB))
        if i % 16 == 15:
            lines += ["meminfo", "nvmlinfo", "ledger 0"]
    lines += ["free 3", "free 100", "ledger 0", "meminfo", "nvmlinfo2"]
    env = {"CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true"}
    (a, ea), (b, eb) = both("\n".join(lines) + "\n", env)
    assert a == b, "reference:\n%s\nb200:\n%s\n%s" % (a[-3000:], b[-3000:], eb[-2000:])
    assert "ledger -> size 1 [self" in a and "-> 2" in a


def test_device_reset_mid_process_matches_reference_on_real_driver(built):
    """cudaDeviceReset() and carry on: the library's device-resident state dies with the primary context
    and is rebuilt (own footprint re-measured) - every reported number still equals the reference's."""
    lines = ["init 0", "alloc %d" % GiB, "alloc %d" % GiB, "meminfo", "nvmlinfo", "reset", "meminfo", "nvmlinfo",
             "alloc %d" % (2 * GiB), "alloc %d" % GiB, "alloc %d" % GiB, "meminfo", "nvmlinfo", "nvmlinfo2", "reset",
             "alloc %d" % (512 * MiB), "meminfo", "nvmlinfo"]
    for args in ((), ("--gpa",)):
        (a, ea), (b, eb) = both("\n".join(lines) + "\n", {"CUDA_MEM_LIMIT_0": "4g", "CUDA_CORE_LIMIT_0": "50"}, args=args)
        assert a == b, "reference:\n%s\nb200:\n%s\n%s" % (a, b, eb[-2000:])
        assert a.count("reset 0 0 0") == 2 and "-> 2" in a


def gpu_uuids():
    try:  # evaluated at collection time, also on machines without a driver
        out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    except OSError:
        return []
    return [l.strip() for l in out.stdout.splitlines() if l.strip()]


@pytest.mark.skipif(len(gpu_uuids()) < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_one_process_on_two_real_gpus_follows_host_indexes(built):
    """Host index 0 = CUDA device 1, host index 1 = a hole, host index 2 = CUDA device 0 (the visible list is
    in host order, CUDA order comes from CUDA_VISIBLE_DEVICES): caps and reported numbers per device must
    match the reference's, which resolves the mapping through the UUIDs like this library does."""
    u = gpu_uuids()
    hole = "GPU-00000000-0000-0000-0000-000000000000"
    env = {"MANAGER_VISIBLE_DEVICES": ",".join((u[1], hole, u[0])), "CUDA_VISIBLE_DEVICES": "0,1",
           "CUDA_MEM_LIMIT_0": "2g", "CUDA_MEM_LIMIT_2": "6g", "CUDA_CORE_LIMIT_0": "30", "CUDA_CORE_LIMIT_2": "60"}
    lines = ["init 0", "totalmem", "meminfo", "alloc %d" % (2 * GiB), "alloc %d" % (2 * GiB), "meminfo", "nvmlinfo",
             "dev 1", "totalmem", "meminfo", "alloc %d" % GiB, "alloc %d" % GiB, "meminfo", "nvmlinfo",
             "dev 0", "meminfo", "alloc %d" % (3 * GiB), "meminfo", "nvmlinfo2"]
    (a, ea), (b, eb) = both("\n".join(lines) + "\n", env)
    assert a == b, "reference:\n%s\nb200:\n%s\n%s" % (a, b, eb[-2000:])
    assert a.count("-> 2") == 2 and str(6 * GiB) in a.splitlines()[1]


def test_launch_storm_under_core_cap_completes_and_is_gated_on_device(built):
    sb = H.Sandbox()
    env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(),
                                      "CUDA_CORE_LIMIT_0": "25", "CUDA_MEM_LIMIT_0": "4g", "CUDA_VISIBLE_DEVICES": "0",
                                      "LOGGER_LEVEL": "1"}, stub=False)
    r = subprocess.run([H.STORM, "--steps", "2", "--warmup", "1", "--per-step", "100000"], env=env, capture_output=True,
                       text=True, timeout=300)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["launches"] == 200000 and d["fails"] == 0
    assert d["sampler_launches"] > 0          # the on-device sampler/controller ran
    assert d["p50_ns"] < 20000                 # the hook never sleeps on the CPU


def test_two_processes_in_one_container_see_reference_numbers(built):
    """Multi-process container: process A holds memory while process B queries and allocates.
    Every library instance owns a few MiB of HBM; the shared own-footprint registry next to the
    GPU lock file must remove A's share from B's view too, or B would differ from the reference."""
    import time
    env = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "LOGGER_LEVEL": "1",
           "CUDA_VISIBLE_DEVICES": "0", "CUDA_MEM_LIMIT_0": "6g"}
    a_script = "init 0\nalloc %d\nalloc %d\nsleepms 9000\nmeminfo\n" % (512 * MiB, 256 * MiB)
    b_script = "init 0\nmeminfo\nnvmlinfo\nalloc %d\nmeminfo\nnvmlinfo2\nalloc %d\nmeminfo\n" % (1 * GiB, 4 * GiB)
    views = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        pa = subprocess.Popen([H.SCENARIO], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=H.preload_env(lib, sb, env, stub=False))
        pa.stdin.write(a_script)
        pa.stdin.flush()
        time.sleep(4.0)  # A has its context and allocations by now
        out_b, err_b, _ = H.run_scenario(lib, b_script, env, sb=sb, stub=False, timeout=120)
        out_a, err_a = pa.communicate(timeout=60)
        sb.cleanup()
        views.append((out_b, out_a, err_b))
    assert views[0][0] == views[1][0], "B's view differs:\nreference:\n%s\nb200:\n%s\n%s" % (views[0][0], views[1][0], views[1][2][-1500:])
    assert views[0][1] == views[1][1]
    assert "-> 2" in views[0][0]  # the 4 GiB request exceeds what is left of the 6 GiB cap


@pytest.mark.parametrize("mode,threads", [("created", 1), ("ptsz", 1), ("created", 4)])
def test_launch_storm_on_other_stream_kinds(built, mode, threads):
    """Created (non-blocking) streams and the per-thread default stream (_ptsz entry points):
    slots, completion markers and device-side gates must work there too, from several threads."""
    sb = H.Sandbox()
    env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(),
                                      "CUDA_CORE_LIMIT_0": "25", "CUDA_MEM_LIMIT_0": "4g", "CUDA_VISIBLE_DEVICES": "0",
                                      "LOGGER_LEVEL": "1"}, stub=False)
    r = subprocess.run([H.STORM, "--steps", "3", "--warmup", "1", "--per-step", "120000", "--stream", mode, "--threads",
                        str(threads), "--max-seconds", "40"], env=env, capture_output=True, text=True, timeout=200)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["fails"] == 0 and d["launches"] >= 120000 and d["limiter"]["present"] == 1
    assert d["sampler_launches"] > 0 and d["p50_ns"] < 20000


@pytest.mark.parametrize("call", ["sync", "htod", "dtoh", "copy"])
def test_blocking_calls_of_a_throttled_tenant_never_need_a_loan(built, call):
    """A 10 %-capped busy tenant walks into a blocking driver call every 100 launches while its stream is
    parked behind the gate.  The reference's thread would be asleep inside the launch hook at that point;
    here the hooked blocking calls (cuCtxSynchronize, the synchronous copies) wait in user space until
    nothing is parked, so the tick thread is never kept out of the driver: every launch completes, the
    cap did bite, and the driver-free watchdog never had to lend tokens."""
    sb = H.Sandbox()
    env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(),
                                      "CUDA_CORE_LIMIT_0": "10", "CUDA_MEM_LIMIT_0": "4g", "CUDA_VISIBLE_DEVICES": "0",
                                      "LOGGER_LEVEL": "1"}, stub=False)
    r = subprocess.run([H.STORM, "--steps", "1000000", "--warmup", "0", "--per-step", "200", "--max-seconds", "5",
                        "--spin-iters", "20000", "--grid", "592", "--block", "256", "--sync-every", "100", "--block-with", call],
                       env=env, capture_output=True, text=True, timeout=300)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["fails"] == 0 and d["launches"] > 0 and d["limiter"]["present"] == 1, d
    assert d["gated_launches"] > 0, d
    assert d["watchdog_loans"] == 0, d


def test_client_mode_registration_and_pids_file_on_real_driver(built):
    """Compatibility mode 200: at start-up the library fork/execs registry/device-client
    (register.c:14-38) and afterwards only pids listed in pids.config count as the container's.  The
    stand-in client registers its parent - the tenant - the way the real one asks the device plugin to;
    memory numbers and the OOM point must match the reference on the real driver."""
    def prep(sb):
        os.makedirs(sb.path("etc/vgpu-manager/registry"), exist_ok=True)
        client = sb.path("etc/vgpu-manager/registry/device-client")
        with open(client, "w") as f:
            f.write("#!/bin/sh\necho $PPID > /etc/vgpu-manager/config/pids.config\nexit 0\n")
        os.chmod(client, 0o755)

    lines = ["init 0", "totalmem", "meminfo", "nvmlinfo", "alloc %d" % GiB, "alloc %d" % GiB, "meminfo", "nvmlinfo",
             "alloc %d" % (3 * GiB), "alloc %d" % (512 * MiB), "meminfo", "nvmlinfo2", "free 0", "meminfo"]
    env = {"MANAGER_COMPATIBILITY_MODE": "200", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "LOGGER_LEVEL": "1", "CUDA_VISIBLE_DEVICES": "0",
           "CUDA_MEM_LIMIT_0": "4g", "CUDA_CORE_LIMIT_0": "50", "VGPU_POD_UID": "uid-1", "VGPU_CONTAINER_NAME": "c",
           "MANAGER_CLIENT_REGISTER_UUID": "r"}
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        sb = H.Sandbox()
        prep(sb)
        out, err, _ = H.run_scenario(lib, "\n".join(lines) + "\n", env, sb=sb, stub=False, timeout=300)
        with open(sb.path("etc/vgpu-manager/config/pids.config")) as f:
            registered = f.read().split()
        outs.append((out, err, registered))
        sb.cleanup()
    (a, ea, ra), (b, eb, rb) = outs
    assert a == b, "reference:\n%s\nb200:\n%s\n%s" % (a, b, eb[-2000:])
    assert len(ra) == 1 and len(rb) == 1  # each tenant was registered by its own client child
    assert "alloc %d -> 2" % (3 * GiB) in a
    # the container's usage really is the tenant's own (the registered pid was found in NVML's list)
    used = [int(l.split()[-1]) for l in a.splitlines() if l.startswith("nvmlinfo ->")]
    assert used[1] - used[0] == 2 * GiB
