"""pytest plumbing: markers, one-time builds, and helpers shared by the suites.

CPU suite  (-m "not gpu"): oracle vs the reference's own code (oracle/_ref), host logic of the
library against the stub driver, symbol surface.  GPU suite (-m gpu): the sm_100a kernels through
the C ABI vs the oracle and the committed golden vectors.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))


@pytest.fixture(scope="session")
def built():
    """Build everything the tests need (idempotent, make/mtime driven)."""
    import helpers
    helpers.build_all()
    return helpers
