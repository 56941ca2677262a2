"""GPU (B200): the compute-share limiter must hold the requested core-% within the REFERENCE's own
tolerance.  The reference defines none (SURVEY.md 8a L-tol), so the band was produced by running
the unmodified reference library five times through every shape on a B200 (`python tests/band.py
--impl reference --runs 5`, committed as tests/golden/tolerance_band.json).  Here the B200 library
runs the same shapes; every enforcement metric (achieved rate, NVML utilisation, what a neighbour
keeps, fairness between tenants, GEMM share) has to land inside [min - spread, max + spread] of the
reference's runs.  Bang-bang control is noisy (the reference's own 10 %-cap storm spreads 19 %), so a
shape that misses once is repeated twice and must then be inside both times.
"""
import json
import os

import pytest

import band
import helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(band.BAND_FILE),
                                                   reason="tests/golden/tolerance_band.json not generated yet")]

_ctx = {}
_report = {}


def _run(shape):
    ref = band.load_band()
    tries = []
    got = band.run_shape(shape, H.NEW_SO, _ctx)
    tries.append(got)
    bad = band.check(shape, got, ref)
    if bad:
        again = [band.run_shape(shape, H.NEW_SO, _ctx) for _ in range(2)]
        tries += again
        bad2 = [v for g in again for v in band.check(shape, g, ref)]
        bad = (bad + bad2) if bad2 else []
    _report[shape] = {"runs": tries, "violations": bad}
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "band_b200_test.json"), "w") as f:
        json.dump(_report, f, indent=1)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("shape", band.CHEAP)
def test_storm_and_sharing_shapes_inside_reference_band(built, shape):
    _run(shape)


@pytest.mark.parametrize("shape", ("gemm1", "gemm4"))
def test_config3_gemm_shapes_inside_reference_band(built, shape):
    _run(shape)
