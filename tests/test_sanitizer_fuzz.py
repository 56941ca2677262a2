"""The host shim rebuilt with -fsanitize=undefined -fno-sanitize-recover and driven by the seeded
random tenant scripts of test_differential_fuzz.py (plus device resets and metered CUDA graphs):
any undefined behaviour (misaligned access into the packed contract files, signed overflow in the
accounting, out-of-range shifts ...) aborts the tenant.  AddressSanitizer cannot be combined with a
library that interposes dlsym itself, so memory errors are covered by the stub driver instead,
which unmaps device and pinned memory as soon as the tenant frees it or resets the device."""
import os
import random
import subprocess

import pytest

import helpers as H
import test_differential_fuzz as F

CSRC = os.path.join(H.ROOT, "vgpu_manager_b200", "csrc")
UBSAN_SO = os.path.join(H.BUILD, "ubsan", "libvgpu-control.so")
SRCS = ["boot.c", "hooktab.c", "config.c", "device.c", "memgate.c", "slabmode.c", "limiter.c", "lifecycle.c", "metrics.c", "kernels_image.gen.c"]


@pytest.fixture(scope="module")
def ubsan_lib():
    H.build_all()
    os.makedirs(os.path.dirname(UBSAN_SO), exist_ok=True)
    r = subprocess.run(["gcc", "-D_GNU_SOURCE", "-std=gnu11", "-O1", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined",
                        "-fno-omit-frame-pointer", "-fPIC", "-shared", "-fvisibility=hidden", "-pthread", "-o", UBSAN_SO,
                        *[os.path.join(CSRC, s) for s in SRCS], "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no UBSan runtime for this gcc: " + r.stderr[-300:])
    return UBSAN_SO


def test_random_tenants_under_ubsan(ubsan_lib):
    rng = random.Random(4242)
    for case in range(50):
        script = F.random_script(rng, rng.randrange(10, 50))
        env, prep = F.random_env(rng), None
        if rng.random() < 0.4:
            env, prep = F.random_membership(rng, env)
        if rng.random() < 0.3:
            script = script.replace("meminfo\n", "meminfo\nreset\n", 1)
        if rng.random() < 0.3:
            env["VGPU_B200_SLAB"] = "1"
        if rng.random() < 0.3:
            env["VGPU_B200_GRAPH_LIMIT"] = "1"
            script += "graph 5 10\ngraphlaunch 20\n"
        args = ("--gpa",) if rng.random() < 0.3 else ()
        sb = H.Sandbox()
        if prep:
            prep(sb)
        e = H.preload_env(ubsan_lib, sb, env)
        e["UBSAN_OPTIONS"] = "print_stacktrace=1:halt_on_error=1"
        r = subprocess.run([H.SCENARIO, *args], input=script, capture_output=True, text=True, env=e, timeout=120)
        sb.cleanup()
        assert "runtime error" not in r.stderr and r.returncode in (0, 1), "case %d rc %d env %r\n%s" % (case, r.returncode, env, r.stderr[-3000:])
