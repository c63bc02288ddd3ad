"""Committed regression fixtures (tests/golden/oracle_lm_small.json, made by tests/golden/make_golden.py): small seeded
problems for every camera model. They are oracle outputs, not reference outputs (the Rust crate cannot be built here;
DESIGN.md section 5), so they pin the oracle against silent changes (CPU) and give the GPU path a second, frozen target."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

from apex_solver_b200 import _ffi as F  # noqa: E402
from oracle_backend import OracleContext  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_lm_small.json")))


def check(got, want, unit_tol, iter_tol, final_tol, same_counts):
    assert abs(got["cost0"] - want["cost0"]) <= unit_tol * want["cost0"]
    assert abs(got["matvec_norm"] - want["matvec_norm"]) <= 100 * unit_tol * want["matvec_norm"]
    assert np.allclose(got["matvec_head"], want["matvec_head"], rtol=0, atol=1e3 * unit_tol * want["matvec_norm"])
    assert (got["status"], got["iterations"], got["accepted"]) == (want["status"], want["iterations"], want["accepted"])
    if same_counts:
        assert got["pcg"] == want["pcg"]
    for a, b in zip(got["costs"], want["costs"]):
        assert abs(a - b) <= iter_tol * b
    assert abs(got["final_cost"] - want["final_cost"]) <= final_tol * want["final_cost"]
    assert np.allclose(got["param_norms"], want["param_norms"], rtol=max(final_tol, 1e-9))


@pytest.mark.parametrize("i", range(len(G.CASES)), ids=[c[0] for c in G.CASES])
def test_oracle_reproduces_its_fixtures(i):
    name, kw, variant = G.CASES[i]
    assert GOLD["shape"] == G.SHAPE
    prob = G.problem(kw, 100 + i)
    got = G.record(OracleContext().upload(prob), prob, variant)
    check(got, GOLD["cases"][name], unit_tol=1e-14, iter_tol=1e-12, final_tol=1e-12, same_counts=True)  # same code, same machine class: libm only


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(G.CASES)), ids=[c[0] for c in G.CASES])
def test_gpu_matches_the_fixtures(i):
    from apex_solver_b200.context import GpuContext
    name, kw, variant = G.CASES[i]
    prob = G.problem(kw, 100 + i)
    got = G.record(GpuContext().upload(prob), prob, variant)
    # Unit stages to 1e-12. Along the LM trajectory the summation-order noise is amplified by the conditioning of the reduced
    # system (measured rounding floor): ~1e-9 with fixed intrinsics, ~1e-6 when the intrinsics are free
    # (the scale gauge is then held by lambda alone); the accept pattern, iteration count and status must be identical.
    tol = 1e-5 if kw["self_calibration"] else 1e-8
    check(got, GOLD["cases"][name], unit_tol=1e-12, iter_tol=tol, final_tol=tol, same_counts=False)
