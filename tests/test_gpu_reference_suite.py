"""GPU (B200): the reference's OWN smoke tests (library/test/test_*.c, test_runtime_launch.cu),
compiled from /root/reference by oracle/Makefile into oracle/_ref/ref_tests/, run the way the
reference's run_all_tests.sh runs them - LD_PRELOAD=<lib> ./test_x, pass = exit 0 within 120 s -
once under the reference library and once under the B200 library, under an active cap.  The
B200 library must exit with the reference's code for every test (one of them, test_alloc_managed,
fails under BOTH libraries with this 8 GiB cap: managed allocations are ledgered and hit it).
For the deterministic single-threaded memory tests the printed NVML usage figures must also be
identical between the two libraries.
"""
import glob
import json
import os
import subprocess

import pytest

import helpers as H

REF_TESTS = os.path.join(H.ROOT, "oracle", "_ref", "ref_tests")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="oracle/_ref/ref_tests did not travel")]
SAME_STDOUT = {"test_alloc", "test_alloc_pitch", "test_alloc_managed", "test_create_array", "test_create_3d_array",
               "test_runtime_alloc", "test_runtime_alloc_managed"}


def gpu0_uuid():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return out.stdout.splitlines()[0].strip()


def run_under(lib, exe):
    sb = H.Sandbox()
    env = H.preload_env(lib, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(),
                                  "CUDA_VISIBLE_DEVICES": "0", "CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "2",
                                  "VMEMORY_NODE_ENABLED": "true", "CUDA_CORE_LIMIT_0": "50", "LOGGER_LEVEL": "1"}, stub=False)
    try:
        r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
        rc, out, err = r.returncode, r.stdout, r.stderr
    except subprocess.TimeoutExpired:
        rc, out, err = 124, "", "timeout"
    sb.cleanup()
    return rc, out, err


def test_reference_smoke_suite_passes_under_both_libraries(built):
    exes = sorted(p for p in glob.glob(os.path.join(REF_TESTS, "test_*")) if os.access(p, os.X_OK))
    assert len(exes) >= 12
    report, failures = {}, []
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    for exe in exes:
        name = os.path.basename(exe)
        rc_ref, out_ref, err_ref = run_under(H.REF_SO, exe)
        rc_new, out_new, err_new = run_under(H.NEW_SO, exe)
        same = out_ref == out_new
        report[name] = {"rc_reference": rc_ref, "rc_b200": rc_new, "stdout_identical": same}
        if rc_new != rc_ref or (rc_ref == 0 and rc_new != 0):
            failures.append((name, rc_ref, rc_new, err_new[-800:]))
        elif name in SAME_STDOUT and not same:
            import difflib
            diff = "\n".join(list(difflib.unified_diff(out_ref.splitlines(), out_new.splitlines(), "reference", "b200", lineterm=""))[:40])
            with open(os.path.join(H.ROOT, "gpurun_out", "refsuite_%s.diff" % name), "w") as f:
                f.write(diff + "\n---- stderr b200\n" + err_new[-3000:])
            failures.append((name, "stdout differs", diff[:1500]))
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "reference_suite_r1.json"), "w") as f:
        json.dump(report, f, indent=1)
    assert not failures, failures
