"""kwta.KWTA / kwta.NeighInhib mirror (emer/vision v1.1.15, with leabra v1.1.48 fffb.Params and nxx1.Params) for the
step SndEnv.ApplyGabor runs after agabor.Convolve (reference sound/sndenv.go:303-323, 481-497).  Third-party packages
that are not in the reference tree: the GPU operator (aud_apply_kwta) restates the published FFFB / noisy-XX1 equations,
parity unpinned -- see include/auditory_b200.h."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib


@dataclass
class FFFBParams:
    """leabra fffb.Params with its Defaults()."""
    On: bool = True
    Gi: float = 1.8
    FF: float = 1.0
    FB: float = 1.0
    FBTau: float = 1.4
    MaxVsAvg: float = 0.0
    FF0: float = 0.1


@dataclass
class NXX1Params:
    """leabra nxx1.Params with its Defaults()."""
    Thr: float = 0.5
    Gain: float = 100.0
    NVar: float = 0.005
    VmActThr: float = 0.01
    SigMult: float = 0.33
    SigMultPow: float = 0.8
    SigGain: float = 3.0
    InterpRange: float = 0.01
    GainCorRange: float = 10.0
    GainCor: float = 0.1


@dataclass
class NeighInhib:
    """vision kwta.NeighInhib."""
    On: bool = False
    Gi: float = 0.6

    def Defaults(self) -> None:
        self.On = True
        self.Gi = 0.6


@dataclass
class KWTA:
    """vision kwta.KWTA."""
    On: bool = False
    Iters: int = 20
    DelActThr: float = 0.005
    LayFFFB: FFFBParams = field(default_factory=FFFBParams)
    PoolFFFB: FFFBParams = field(default_factory=FFFBParams)
    XX1: NXX1Params = field(default_factory=NXX1Params)
    ActTau: float = 3.0
    Gbar: Tuple[float, float, float, float] = (0.5, 0.1, 1.0, 1.0)     # E, L, I, K
    Erev: Tuple[float, float, float, float] = (1.0, 0.3, 0.25, 0.25)

    def Defaults(self) -> None:
        self.On = True
        self.Iters = 20
        self.DelActThr = 0.005
        self.LayFFFB, self.PoolFFFB = FFFBParams(), FFFBParams()
        self.PoolFFFB.Gi = 2.0
        self.XX1 = NXX1Params()
        self.XX1.Gain = 80.0
        self.XX1.NVar = 0.01
        self.ActTau = 3.0
        self.Gbar = (0.5, 0.1, 1.0, 1.0)
        self.Erev = (1.0, 0.3, 0.25, 0.25)


def _fffb(p: FFFBParams) -> _lib.AudFffbParams:
    return _lib.AudFffbParams(int(p.On), p.Gi, p.FF, p.FB, p.FBTau, p.MaxVsAvg, p.FF0)


def params(k: KWTA, ni: NeighInhib, pool_mode: bool) -> _lib.AudKwtaParams:
    x = k.XX1
    return _lib.AudKwtaParams(int(k.On), int(k.Iters), k.DelActThr, _fffb(k.LayFFFB), _fffb(k.PoolFFFB),
                              x.Thr, x.Gain, x.NVar, x.VmActThr, x.SigMult, x.SigMultPow, x.SigGain, x.InterpRange,
                              x.GainCorRange, x.GainCor, k.ActTau, *k.Gbar, *k.Erev, int(pool_mode), int(ni.On), ni.Gi)


def Apply(k: KWTA, ni: NeighInhib, pool_mode: bool, gabor: np.ndarray, shape: Sequence[int],
          seq_base: Optional[np.ndarray] = None, device: int = 0):
    """ApplyNeighInhib + ApplyKwta on a stack of GborOutput tensors [n][prod(shape)]; returns (GborKwta, ExtGi)
    stacks.  seq_base: runs of tensors one SndEnv would produce in order (KWTAPool keeps per-pool state across calls)."""
    g = np.ascontiguousarray(gabor, dtype=np.float32).reshape(len(gabor), -1)
    shp = np.ascontiguousarray(shape, dtype=np.int32)
    out, ext = np.zeros_like(g), np.zeros_like(g)
    sb = None if seq_base is None else np.ascontiguousarray(seq_base, dtype=np.int64)
    kp = params(k, ni, pool_mode)
    _lib.check(_lib.lib().aud_apply_kwta(device, C.byref(kp), g.ctypes.data, g.shape[0], len(shp), shp.ctypes.data,
                                         None if sb is None else sb.ctypes.data, 0 if sb is None else len(sb) - 1,
                                         ext.ctypes.data, out.ctypes.data))
    return out, ext
