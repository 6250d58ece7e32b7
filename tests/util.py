"""Shared helpers for the parity tests."""
import numpy as np

# BASELINE.json north_star tolerances: GPU float32 vs float64 oracle.
# "relative" is taken against max(1, |ref|): log-mel / log-power values sit
# around 0 (ln of sums near 1), where a pure relative error is meaningless.
RTOL_LOG = 1e-4     # log-power, log-mel, energy, mfcc
RTOL_GABOR = 1e-3   # gabor outputs


def assert_close(got, ref, rtol, name=""):
    """|got - ref| <= rtol * max(1, |ref|) on every finite reference value; where the reference is NaN the
    result must be NaN, where it is +-Inf the result must be the same infinity (a NaN never passes as close)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64).reshape(got.shape)
    fin = np.isfinite(ref)
    same_special = np.where(np.isnan(ref), np.isnan(got), got == ref)
    if not same_special[~fin].all():
        raise AssertionError(f"{name}: {(~same_special[~fin]).sum()} non-finite reference values are not reproduced")
    err = np.where(fin, np.abs(got - np.where(fin, ref, 0.0)), 0.0)
    tol = rtol * np.maximum(1.0, np.abs(np.where(fin, ref, 0.0)))
    bad = ~(err <= tol)          # a NaN result against a finite reference is bad
    if bad.any():
        e2 = np.where(np.isnan(err), np.inf, err)
        i = np.unravel_index(np.argmax(e2 / tol), err.shape)
        raise AssertionError(f"{name}: {bad.sum()} of {bad.size} values outside {rtol:g} x max(1,|ref|); "
                             f"worst at {i}: got {got[i]!r} ref {ref[i]!r}")
    worst = float((err / np.maximum(1.0, np.abs(np.where(fin, ref, 0.0)))).max()) if err.size else 0.0
    record(name or "unnamed", worst, rtol)
    return worst


# Margins as evidence: every assert_close records its worst error / max(1,|ref|); the GPU session writes them
# to gpurun_out/parity_errors.json (conftest.py), and the round's copy is committed under profiles/.
RECORDED = {}


def record(name, worst, rtol):
    import os
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    key = f"{test}::{name}"
    prev = RECORDED.get(key)
    if prev is None or worst > prev["max_err_over_max1ref"]:
        RECORDED[key] = {"max_err_over_max1ref": worst, "tolerance": rtol}
