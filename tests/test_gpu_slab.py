"""GPU (B200): VGPU_B200_SLAB=1 - the spill path with real data movement - on the REAL driver.

(1) Accounting: BASELINE config 4 (8 GiB virtual over 2 GiB physical, 64 MiB allocations until the
    cap refuses) - the transcript (return codes, cuMemGetInfo, NVML view, ledger bytes) must be the
    reference's, byte for byte, although every "UVA" decision now demotes an HBM slab to host
    memory with vgpu_spill_copy_kernel instead of calling cuMemAllocManaged.
(2) Integrity + the numbers the hooks measured while doing it: every buffer keeps its contents
    through demotion and promotion; spill / scrub bandwidth from CUDA events around the kernels the
    cuMemAlloc hook launched (gpurun_out/slab_spill_r2.json).
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


def env4():
    """BASELINE config 4's shape: a cap oversold 4x with ~2 GiB of physical head-room.  HOST compatibility mode
    counts every process on the GPU as the container's - including this pytest process when earlier tests
    left a CUDA context (and torch's cached blocks) in it - so the physical share is sized on top of what
    is in use right now: real = used_now + 2 GiB, cap = 4 x real."""
    out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=memory.used", "--format=csv,noheader,nounits"],
                         capture_output=True, text=True).stdout.split()
    used_mib = int(out[0]) if out else 0
    return {"CUDA_MEM_LIMIT_0": "%dm" % (4 * (used_mib + 2048)), "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true"}


def gpu0_uuid():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return out.stdout.splitlines()[0].strip()


def run(lib, script, env):
    base = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "LOGGER_LEVEL": "1", "CUDA_VISIBLE_DEVICES": "0"}
    base.update(env)
    sb = H.Sandbox()
    out, err, _ = H.run_scenario(lib, script, base, sb=sb, stub=False, timeout=600)
    sb.cleanup()
    return out, err


def test_slab_mode_accounting_matches_reference_on_real_driver(built):
    lines = ["init 0", "nvmlinit 0"]
    for i in range(132):
        lines.append("alloc %d" % (64 * MiB))
        if i % 16 == 15:
            lines += ["meminfo", "nvmlinfo", "ledger 0"]
    lines += ["free 3", "free 100", "ledger 0", "meminfo", "nvmlinfo2", "free 40", "free 41", "nvmlinfo", "ledger 0",
              "alloc %d" % (64 * MiB), "alloc %d" % (32 * MiB), "nvmlinfo", "ledger 0"]
    script = "\n".join(lines) + "\n"
    ENV4 = env4()
    ref, _ = run(H.REF_SO, script, ENV4)
    slab, err = run(H.NEW_SO, script + "slabstats 0\n", dict(ENV4, VGPU_B200_SLAB="1"))
    body = "\n".join(slab.splitlines()[:-1]) + "\n"
    assert body == ref, "reference:\n%s\nslab mode:\n%s\n%s" % (ref[-3000:], body[-3000:], err[-3000:])
    st = slab.splitlines()[-1].split()
    st = dict(zip(st[1::2], map(int, st[2::2])))
    assert "slab mode disabled" not in err, err[-2000:]
    assert st["allocs"] > 100 and st["demotions"] > 10 and st["spill_bytes"] == st["demotions"] * 64 * MiB, st


def test_slab_mode_moves_data_and_keeps_it_intact(built):
    n = 256 * MiB
    lines = ["init 0", "nvmlinit 0"]
    vals = {}
    k = 0
    for i in range(12):  # 2 GiB physical: the later ones spill the earliest
        lines += ["alloc %d" % n, "fill %d %d %d" % (i, n, 11 + 7 * i)]
        vals[i] = 11 + 7 * i
    lines += ["nvmlinfo", "ledger 0"]
    for i in range(12):
        lines.append("check %d %d %d" % (i, n, vals[i]))
    lines += ["free 11", "free 0", "free 5"]
    for i in (1, 2, 3, 4, 6, 7, 8, 9, 10):
        lines.append("check %d %d %d" % (i, n, vals[i]))
    lines += ["nvmlinfo", "ledger 0", "slabstats 0"]
    ENV4 = env4()
    out, err = run(H.NEW_SO, "\n".join(lines) + "\n", dict(ENV4, VGPU_B200_SLAB="1"))
    assert "slab mode disabled" not in err, err[-2000:]
    assert "CORRUPT" not in out and out.count("intact") == 21, out[-3000:] + err[-2000:]
    st = out.splitlines()[-1].split()
    st = dict(zip(st[1::2], map(int, st[2::2])))
    assert st["demotions"] >= 4 and st["spill_bytes"] == st["demotions"] * n and st["scrubbed_bytes"] >= 4 * n, st
    # what the hooks measured (CUDA events around the kernels they launched): metric ids from vgpu_internal.h
    probe = ["init 0", "nvmlinit 0"] + ["alloc %d" % GiB for _ in range(5)] + ["slabstats 0"]
    out2, err2 = run(H.NEW_SO, "\n".join(probe) + "\n", dict(ENV4, VGPU_B200_SLAB="1", LOGGER_LEVEL="3"))
    report = {"integrity": st, "one_gib_slabs": out2.splitlines()[-1], "log_tail": err2[-1500:]}
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "slab_spill_r2.json"), "w") as f:
        json.dump(report, f, indent=1)
