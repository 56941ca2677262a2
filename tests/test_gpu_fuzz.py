"""GPU (B200): the seeded random differential test of tests/test_differential_fuzz.py on the REAL
driver - random allocation / free / report scripts under random caps, run under the compiled
reference and under the B200 library; transcripts (return codes and every reported number) must be
identical.  On real hardware `used` includes the CUDA context and, for the B200 library, its own
HBM block + module, which the quota kernel has to remove exactly for this to pass."""
import os
import random

import pytest

import helpers as H
import test_differential_fuzz as F
from test_gpu_differential import gpu0_uuid

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref/libvgpu-control.so did not travel")]
MiB, GiB = 1 << 20, 1 << 30


def gpu_env(rng):
    env = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "LOGGER_LEVEL": "0",
           "CUDA_VISIBLE_DEVICES": "0"}
    env["CUDA_MEM_LIMIT_0"] = rng.choice(("2g", "4g", "6g", "1536m"))
    if rng.random() < 0.5:
        env["CUDA_MEM_RATIO_0"] = rng.choice(("2", "4"))
    if rng.random() < 0.6:
        env["VMEMORY_NODE_ENABLED"] = "true"
    if rng.random() < 0.3:
        env["CUDA_CORE_LIMIT_0"] = rng.choice(("30", "100"))
    return env


def run(lib, script, env):
    sb = H.Sandbox()
    out, err, rc = H.run_scenario(lib, script, env, sb=sb, stub=False, check=False, timeout=300)
    sb.cleanup()
    return out, rc, err


def test_random_scripts_on_the_real_driver(built):
    rng = random.Random(0x5EED)
    for case in range(14):
        script = F.random_script(rng, rng.randrange(10, 45))
        env = gpu_env(rng)
        ref = run(H.REF_SO, script, env)
        new = run(H.NEW_SO, script, env)
        assert ref[:2] == new[:2], "case %d env %r\nscript:\n%s\n--- reference (rc %d)\n%s\n--- b200 (rc %d)\n%s\n%s" % (
            case, env, script, ref[1], ref[0], new[1], new[0], new[2][-1500:])
