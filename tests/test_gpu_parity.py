"""GPU parity: fused sm_100a kernel (through the C-ABI) vs the float64 oracle."""
import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import synth
from oracle import c_oracle
from util import RTOL_GABOR, RTOL_LOG, assert_close

pytestmark = pytest.mark.gpu


def make_env(sig, mfcc=True, deltas=True, gabor=True, prev=0.0, cur=None, out4d=True, by_time=False):
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(sig, synth.SR)
    se.Mel.MFCC = mfcc
    se.Mel.Deltas = deltas
    if gabor:
        synth.configure_processspeech_gabor(se, out4d=out4d, by_time=by_time)
    se.Init()
    se.DFT.PrevSmooth = prev
    se.DFT.CurSmooth = (1.0 - prev) if cur is None else cur
    return se


def oracle_env(mfcc=True, deltas=True, gabor=True, prev=0.0, cur=None, out4d=True, by_time=False):
    p = c_oracle.default_params(mfcc=int(mfcc), deltas=int(deltas), prev_smooth=prev,
                                cur_smooth=(1.0 - prev) if cur is None else cur)
    specs = []
    if gabor:
        c_oracle.with_processspeech_gabor(p, out4d=out4d, by_time=by_time)
        specs = c_oracle.processspeech_specs()
    return c_oracle.Env(p, specs)


def compare(got, ref, names):
    worst = {}
    for n in names:
        rtol = RTOL_GABOR if n == "gabor" else RTOL_LOG
        if n in ("power",):
            # raw power spans 8 decades; compare relative to the frame's peak as well
            g, r = got[n].astype(np.float64), ref[n].reshape(got[n].shape)
            scale = np.maximum(np.abs(r).max(axis=1, keepdims=True), 1.0)
            assert np.all(np.abs(g - r) <= 2e-5 * scale), "power"
            continue
        if n in ("deltas", "delta_deltas"):
            g, r = got[n].astype(np.float64), ref[n].reshape(got[n].shape)
            scale = max(1.0, np.abs(r).max())
            assert np.abs(g - r).max() <= RTOL_LOG * scale, n
            continue
        worst[n] = assert_close(got[n], ref[n], rtol, n)
    return worst


@pytest.mark.parametrize("prev", [0.0, 0.3])
def test_config1_all_outputs(prev):
    sig = synth.config1_signal()
    se = make_env(sig, prev=prev)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "energy", "mfcc", "deltas", "delta_deltas", "gabor",
                                                      "power", "logpower"])
    ref = oracle_env(prev=prev).process(sig.astype(np.float64), want_power=True)
    assert got["mel"].shape == (20, 32, 14)
    assert got["gabor"].shape == (20, 256)
    w = compare(got, ref, ["mel", "energy", "mfcc", "deltas", "delta_deltas", "gabor", "power", "logpower"])
    print(w)
    # tail: last segment's steps 10..13 run past the signal -> exact zeros (sndenv.go:355-358)
    assert np.all(got["mel"][19, :, 10:] == 0.0)
    assert np.all(ref["mel"][19, :, 10:] == 0.0)
