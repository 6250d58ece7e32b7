"""
np_oracle.py -- float64 numpy restatement of emer/auditory's speech-feature path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and only as the checker.  The product path (auditory_b200/)
never imports, links or executes anything in this directory.

PARITY UNPINNED.  The reference (Go, v0.9.8) ships no tests, golden vectors or
fixtures for this path, and there is no Go toolchain in the build container, so
the reference itself cannot be run to pin this restatement.  The substitutes
are (i) analytic known-answer tests (tests/test_oracle_kat.py), (ii) the
BinPts table SURVEY.md section 8(a6) derived from the Go formulas, and (iii) a
cross-check between this file and the independent C twin
(oracle/auditory_oracle.c).  Third-party arithmetic the Go code calls and that
is not vendored under /root/reference:
  gonum.org/v1/gonum v0.11.0 dsp/fourier (go.mod:20)
     CmplxFFT.Coefficients = forward, unnormalised DFT, sign e^{-2 pi i jk/n}
        (FFTPACK cfftf)                         -> numpy.fft.fft
     DCT.Transform = FFTPACK cost = unnormalised DCT-I
        y[k] = x[0] + (-1)^k x[n-1] + 2 sum_{j=1}^{n-2} x[j] cos(pi j k/(n-1))
                                                -> dct1() below (== scipy dct type 1)
  github.com/emer/etable v1.1.7 etensor: row-major dense tensors whose
     Set/Value([]int) use stride arithmetic with no per-dimension bounds check
     and whose FloatValRowCell(row, cell) = Values[row*(Len/Dim0)+cell].

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


# --------------------------------------------------------------------------
# sound/sndenv.go
# --------------------------------------------------------------------------
def go_round(x: float) -> float:
    """Go math.Round: half away from zero (Python's round() is half-to-even)."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def msec_to_samples(ms: float, rate: int) -> int:
    """sound/sndenv.go:522-524 MSecToSamples."""
    return int(go_round(ms * 0.001 * float(rate)))


@dataclass
class SoundParams:
    """sound/sndenv.go:24-61 Params + :64-71 ParamDefaults."""
    WinMs: float = 25.0
    StepMs: float = 10.0
    SegmentMs: float = 100.0
    StrideMs: float = 100.0
    BorderSteps: int = 2
    Channel: int = 0
    WinSamples: int = 0
    StepSamples: int = 0
    SegmentSamples: int = 0
    StrideSamples: int = 0
    SegmentSteps: int = 0
    Steps: List[int] = field(default_factory=list)


# --------------------------------------------------------------------------
# dft/dft.go
# --------------------------------------------------------------------------
@dataclass
class DftParams:
    """dft/dft.go:15-39 Params + Defaults."""
    CompLogPow: bool = True
    LogMin: float = -100.0
    LogOffSet: float = 1.0
    PrevSmooth: float = 0.0
    CurSmooth: float = 1.0

    def Defaults(self):
        self.PrevSmooth = 0.0
        self.CurSmooth = 1.0 - self.PrevSmooth
        self.CompLogPow = True
        self.LogOffSet = 1.0
        self.LogMin = -100.0

    def Filter(self, step, window, win_samples, power, log_power, power_seg, log_power_seg):
        """dft/dft.go:42-50: complex copy (FftReal :53-59), forward FFT of
        length win_samples (gonum CmplxFFT.Coefficients), then Power."""
        coefs = np.fft.fft(np.asarray(window[:win_samples], dtype=np.float64).astype(np.complex128))
        self.Power(step, win_samples, coefs, power, log_power, power_seg, log_power_seg)

    def Power(self, step, win_samples, coefs, power, log_power, power_seg, log_power_seg):
        """dft/dft.go:62-85."""
        for k in range(win_samples // 2 + 1):
            rl = coefs[k].real
            im = coefs[k].imag
            powr = rl * rl + im * im
            if step > 0:
                powr = self.PrevSmooth * power[k] + self.CurSmooth * powr
            power[k] = powr
            power_seg[k, step] = powr
            if self.CompLogPow:
                powr += self.LogOffSet
                if powr == 0:
                    logp = self.LogMin
                else:
                    logp = math.log(powr)
                log_power[k] = logp
                log_power_seg[k, step] = logp


# --------------------------------------------------------------------------
# mel/mel.go
# --------------------------------------------------------------------------
def freq_to_mel(freq: float) -> float:
    """mel/mel.go:156-158."""
    return 1127.0 * math.log(1.0 + freq / 700.0)


def mel_to_freq(mel: float) -> float:
    """mel/mel.go:161-163."""
    return 700.0 * (math.exp(mel / 1127.0) - 1.0)


def freq_to_bin(freq: float, n_fft: float, sample_rate: float) -> int:
    """mel/mel.go:166-168."""
    return int(math.floor(((n_fft + 1) * freq) / sample_rate))


def dct1(x: np.ndarray) -> np.ndarray:
    """gonum fourier.DCT.Transform == FFTPACK cost: unnormalised DCT-I."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    k = np.arange(n)
    out = x[0] + np.where(k % 2 == 0, 1.0, -1.0) * x[n - 1]
    if n > 2:
        j = np.arange(1, n - 1)
        out = out + 2.0 * (np.cos(np.pi * np.outer(k, j) / (n - 1)) @ x[1:n - 1])
    return out


@dataclass
class MelFilterBank:
    """mel/mel.go:16-44 FilterBank + :171-180 Defaults."""
    NFilters: int = 32
    LoHz: float = 0.0
    HiHz: float = 8000.0
    LogOff: float = 0.0
    LogMin: float = -10.0
    Renorm: bool = True
    RenormMin: float = -6.0
    RenormMax: float = 4.0
    RenormScale: float = 0.0


@dataclass
class MelParams:
    """mel/mel.go:47-74 Params + Defaults (MFCC and Deltas default ON, F8)."""
    FBank: MelFilterBank = field(default_factory=MelFilterBank)
    BinPts: Optional[np.ndarray] = None
    HzPts: Optional[np.ndarray] = None
    MFCC: bool = True
    Deltas: bool = True
    NCoefs: int = 13

    def InitFilters(self, dft_size: int, sample_rate: int) -> np.ndarray:
        """mel/mel.go:77-117.  Returns the filter table with the reference's
        geometry [NFilters, NFilters+2]; etensor Set([]int{f, fi}) is flat
        stride arithmetic (f*(NFilters+2)+fi) with only the slice bound
        checked, so wide filters spill into the next row and an offset past
        the end panics -- reproduced as IndexError."""
        fb = self.FBank
        nf = fb.NFilters
        self.BinPts = np.zeros(nf + 2, dtype=np.int32)
        self.HzPts = np.zeros(nf + 2, dtype=np.float64)
        fb.Renorm = False                                   # mel.go:80
        hi_mel = freq_to_mel(fb.HiHz)
        lo_mel = freq_to_mel(fb.LoHz)
        incr = (hi_mel - lo_mel) / float(nf + 1)
        for i in range(nf + 2):
            ml = lo_mel + float(i) * incr
            hz = mel_to_freq(ml)
            self.HzPts[i] = hz
            self.BinPts[i] = freq_to_bin(hz, float(dft_size), float(sample_rate))
        max_bins = nf + 2
        flat = np.zeros(nf * max_bins, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            for f in range(nf):
                bin_min = int(self.BinPts[f])
                bin_ctr = int(self.BinPts[f + 1])
                bin_max = int(self.BinPts[f + 2])
                pkmin = np.float64(bin_ctr) - np.float64(bin_min)
                pkmax = np.float64(bin_max) - np.float64(bin_ctr)
                fi = 0
                b = bin_min
                while b <= bin_ctr:
                    off = f * max_bins + fi
                    if off >= flat.size:
                        raise IndexError("mel.InitFilters: index out of range (reference panics)")
                    flat[off] = (np.float64(b) - np.float64(bin_min)) / pkmin
                    b += 1
                    fi += 1
                while b <= bin_max:
                    off = f * max_bins + fi
                    if off >= flat.size:
                        raise IndexError("mel.InitFilters: index out of range (reference panics)")
                    flat[off] = (np.float64(bin_max) - np.float64(b)) / pkmax
                    b += 1
                    fi += 1
        return flat.reshape(nf, max_bins)

    def FilterDft(self, step, power, mel_seg, mel_fbank, filters):
        """mel/mel.go:120-153.  filters is read with flat stride arithmetic."""
        fb = self.FBank
        flat = filters.reshape(-1)
        stride = filters.shape[1]
        for flt in range(fb.NFilters):
            min_bin = int(self.BinPts[flt])
            max_bin = int(self.BinPts[flt + 2])
            s = 0.0
            fi = 0
            for b in range(min_bin, max_bin + 1):
                s += flat[flt * stride + fi] * power[b]
                fi += 1
            s += fb.LogOff
            if s == 0:
                val = fb.LogMin
            else:
                val = math.log(s) if s > 0 else float("nan")
            if fb.Renorm:
                val -= fb.RenormMin
                if val < 0.0:
                    val = 0.0
                val *= fb.RenormScale
                if val > 1.0:
                    val = 1.0
            mel_fbank[flt] = val
            mel_seg[flt, step] = val

    def CepstrumDct(self, step, mel_fbank, mfcc_seg):
        """mel/mel.go:192-212."""
        out = dct1(mel_fbank)
        el0 = out[0]
        out[0] = math.log(1.0 + el0 * el0)
        for i in range(self.NCoefs):
            mfcc_seg[i, step] = out[i]


# --------------------------------------------------------------------------
# emer/vision v1.1.15 kwta + emer/leabra v1.1.48 fffb / nxx1 (go.mod:8-9) -- THIRD PARTY, ABSENT FROM /root/reference.
# PARITY UNPINNED: restated from the published FFFB / NXX1 algorithm and from memory of those versions; the call sites
# are sound/sndenv.go:303-323 (ApplyNeighInhib, ApplyKwta).  All arithmetic is float32, as in the Go packages.
# --------------------------------------------------------------------------
F32 = np.float32


class NXX1Params:
    """leabra nxx1.Params: noisy x/(x+1) rate-code activation with a sigmoidal foot below threshold."""

    def __init__(self):
        self.Thr, self.Gain, self.NVar = F32(0.5), F32(100), F32(0.005)
        self.VmActThr, self.SigMult, self.SigMultPow, self.SigGain = F32(0.01), F32(0.33), F32(0.8), F32(3.0)
        self.InterpRange, self.GainCorRange, self.GainCor = F32(0.01), F32(10.0), F32(0.1)
        self.Update()

    def Update(self):
        self.SigGainNVar = F32(self.SigGain / self.NVar)
        self.SigMultEff = F32(self.SigMult * F32(np.power(F32(self.Gain * self.NVar), self.SigMultPow)))
        self.SigValAt0 = F32(F32(0.5) * self.SigMultEff)
        self.InterpVal = F32(self.XX1GainCor(self.InterpRange) - self.SigValAt0)

    def XX1(self, x):
        x = F32(x * self.Gain)
        return F32(x / F32(x + F32(1)))

    def XX1GainCor(self, x):
        fact = F32(F32(self.GainCorRange - F32(x / self.NVar)) / self.GainCorRange)
        if fact < 0:
            return self.XX1(x)
        new_gain = F32(self.Gain * F32(F32(1) - F32(self.GainCor * fact)))
        x = F32(x * new_gain)
        return F32(x / F32(x + F32(1)))

    def NoisyXX1(self, x):
        x = F32(x)
        if x < 0:
            return F32(self.SigMultEff / F32(F32(1) + F32(np.exp(F32(-F32(x * self.SigGainNVar))))))
        if x < self.InterpRange:
            interp = F32(F32(1) - F32(F32(self.InterpRange - x) / self.InterpRange))
            return F32(self.SigValAt0 + F32(interp * self.InterpVal))
        return self.XX1GainCor(x)


class FFFBParams:
    """leabra fffb.Params."""

    def __init__(self):
        self.On, self.Gi, self.FF, self.FB, self.FBTau, self.MaxVsAvg, self.FF0 = True, F32(1.8), F32(1), F32(1), F32(1.4), F32(0), F32(0.1)
        self.Update()

    def Update(self):
        self.FBDt = F32(F32(1) / self.FBTau)

    def FFInhib(self, avg_ge, max_ge):
        ff_netin = F32(avg_ge + F32(self.MaxVsAvg * F32(max_ge - avg_ge)))
        return F32(self.FF * F32(ff_netin - self.FF0)) if ff_netin > self.FF0 else F32(0)


class FFFBInhib:
    """leabra fffb.Inhib: FFi, FBi, Gi and the Ge / Act average-max accumulators."""

    def __init__(self):
        self.FFi = self.FBi = self.Gi = F32(0)
        self.GeAvg = self.GeMax = self.ActAvg = self.ActMax = F32(0)


def _avg_max(vals):
    """minmax.AvgMax32: Init, UpdateVal over vals, CalcAvg (sequential float32 sum)."""
    acc = F32(0)
    mx = F32(-3.4028235e38)
    for v in vals:
        acc = F32(acc + F32(v))
        if v > mx:
            mx = F32(v)
    n = len(vals)
    return (F32(acc / F32(n)) if n > 0 else acc), mx


class KWTA:
    """emer/vision kwta.KWTA with its Defaults()."""

    def __init__(self):
        self.On, self.Iters, self.DelActThr = True, 20, F32(0.005)
        self.LayFFFB, self.PoolFFFB = FFFBParams(), FFFBParams()
        self.PoolFFFB.Gi = F32(2.0)
        self.XX1 = NXX1Params()
        self.XX1.Gain, self.XX1.NVar = F32(80), F32(0.01)
        self.ActTau = F32(3)
        self.GbarE, self.GbarL, self.GbarI, self.GbarK = F32(0.5), F32(0.1), F32(1.0), F32(1.0)
        self.ErevE, self.ErevL, self.ErevI, self.ErevK = F32(1.0), F32(0.3), F32(0.25), F32(0.25)
        self.Update()

    def Update(self):
        self.LayFFFB.Update()
        self.PoolFFFB.Update()
        self.XX1.Update()
        thr = self.XX1.Thr
        self.ErevSubThrE, self.ErevSubThrL, self.ErevSubThrI = F32(self.ErevE - thr), F32(self.ErevL - thr), F32(self.ErevI - thr)
        self.ThrSubErevE, self.ThrSubErevL, self.ThrSubErevI = F32(thr - self.ErevE), F32(thr - self.ErevL), F32(thr - self.ErevI)
        self.ActDt = F32(F32(1) / self.ActTau)

    def GeThrFmG(self, gi):
        return F32(F32(F32(F32(self.GbarI * gi) * self.ErevSubThrI) + F32(self.GbarL * self.ErevSubThrL)) / self.ThrSubErevE)

    def ActFmG(self, ge_thr, ge, act):
        nw = self.XX1.NoisyXX1(F32(F32(ge * self.GbarE) - ge_thr))
        del_act = F32(self.ActDt * F32(nw - act))
        return F32(act + del_act), del_act

    def _fffb(self, fb: FFFBParams, inh: FFFBInhib):
        if not fb.On:
            inh.FFi = inh.FBi = inh.Gi = F32(0)
            return
        ffi = fb.FFInhib(inh.GeAvg, inh.GeMax)
        fbi = F32(fb.FB * inh.ActAvg)
        inh.FFi = ffi
        inh.FBi = F32(inh.FBi + F32(fb.FBDt * F32(fbi - inh.FBi)))
        inh.Gi = F32(fb.Gi * F32(ffi + inh.FBi))

    def KWTALayer(self, raw: np.ndarray, act: np.ndarray, ext_gi: Optional[np.ndarray]):
        """One level of inhibition over every value of the tensor; act comes in as a copy of raw (sndenv.go:315)."""
        raws, acts = raw.reshape(-1), act.reshape(-1)
        exts = None if ext_gi is None else ext_gi.reshape(-1)
        inh = FFFBInhib()
        inh.GeAvg, inh.GeMax = _avg_max(raws)
        for cy in range(self.Iters):
            self._fffb(self.LayFFFB, inh)
            max_del = F32(0)
            for i in range(acts.size):
                gi = inh.Gi if exts is None else F32(inh.Gi + exts[i])
                nw, dl = self.ActFmG(self.GeThrFmG(gi), raws[i], acts[i])
                max_del = max(max_del, F32(abs(dl)))
                acts[i] = nw
            inh.ActAvg, inh.ActMax = _avg_max(acts)
            if cy > 2 and max_del < self.DelActThr:
                break

    def KWTAPool(self, raw: np.ndarray, act: np.ndarray, inhibs: list, ext_gi: Optional[np.ndarray]):
        """Layer-level inhibition over the pools' averages plus pool-level inhibition inside the inner two dimensions.
        `inhibs` (one FFFBInhib per pool) is the caller's persistent state (SndEnv.Inhibs): FBi carries over from the
        previous call, exactly as the reference's reuse of se.Inhibs does."""
        lay_n = raw.shape[0] * raw.shape[1]
        pl_n = raw.shape[2] * raw.shape[3]
        raws, acts = raw.reshape(lay_n, pl_n), act.reshape(lay_n, pl_n)
        exts = None if ext_gi is None else ext_gi.reshape(lay_n, pl_n)
        if len(inhibs) != lay_n:
            inhibs[:] = [FFFBInhib() for _ in range(lay_n)]
        lay = FFFBInhib()
        for pi in range(lay_n):
            inhibs[pi].GeAvg, inhibs[pi].GeMax = _avg_max(raws[pi])
        lay.GeAvg, lay.GeMax = _avg_max([q.GeAvg for q in inhibs])
        for cy in range(self.Iters):
            self._fffb(self.LayFFFB, lay)
            max_del = F32(0)
            for pi in range(lay_n):
                pl = inhibs[pi]
                self._fffb(self.PoolFFFB, pl)
                gi_pool = max(lay.Gi, pl.Gi)
                for ui in range(pl_n):
                    gi = gi_pool
                    if exts is not None:
                        e_in = exts[pi, ui]
                        gi = max(gi, F32(self.PoolFFFB.Gi * self.PoolFFFB.FFInhib(e_in, e_in)))
                    nw, dl = self.ActFmG(self.GeThrFmG(gi), raws[pi, ui], acts[pi, ui])
                    max_del = max(max_del, F32(abs(dl)))
                    acts[pi, ui] = nw
                pl.ActAvg, pl.ActMax = _avg_max(acts[pi])
            lay.ActAvg, lay.ActMax = _avg_max([q.ActAvg for q in inhibs])
            if cy > 2 and max_del < self.DelActThr:
                break


class NeighInhib:
    """emer/vision kwta.NeighInhib: each unit is inhibited by the same feature at its two neighbours along the
    direction orthogonal to the feature's angle (4 angles; inner-most dimension = angle)."""
    ORTHO4X = (0, -1, 1, -1)
    ORTHO4Y = (1, 1, 0, -1)

    def __init__(self):
        self.On, self.Gi = True, F32(0.6)

    def Inhib4(self, act: np.ndarray, ext_gi: np.ndarray):
        lay_y, lay_x, pl_y, pl_x = act.shape
        ext_gi[...] = 0
        for ly in range(lay_y):
            for lx in range(lay_x):
                for py in range(pl_y):
                    for ang in range(min(4, pl_x)):
                        gi = F32(0)
                        for sgn in (1, -1):
                            nx, ny = lx + sgn * self.ORTHO4X[ang], ly + sgn * self.ORTHO4Y[ang]
                            if 0 <= nx < lay_x and 0 <= ny < lay_y:
                                gi = max(gi, F32(self.Gi * act[ny, nx, py, ang]))
                        ext_gi[ly, lx, py, ang] = gi


# --------------------------------------------------------------------------
# agabor/gabor.go
# --------------------------------------------------------------------------
@dataclass
class GaborFilter:
    """agabor/gabor.go:17-42 Filter."""
    Off: bool = False
    WaveLen: float = 0.0
    Orientation: float = 0.0
    SigmaWidth: float = 0.0
    SigmaLength: float = 0.0
    PhaseOffset: float = 0.0
    CircleEdge: bool = False
    Circular: bool = False

    def Defaults(self):
        """agabor/gabor.go:73-86 (prints elided)."""
        if self.WaveLen == 0:
            self.WaveLen = 2
        if self.SigmaLength == 0 and not self.Circular:
            self.SigmaLength = 0.5
        if self.SigmaWidth == 0:
            self.SigmaWidth = 0.5


@dataclass
class GaborFilterSet:
    """agabor/gabor.go:45-70 FilterSet."""
    SizeX: int = 0
    SizeY: int = 0
    StrideX: int = 0
    StrideY: int = 0
    Gain: float = 0.0
    Distribute: bool = False
    Filters: Optional[np.ndarray] = None      # [nf, SizeY, SizeX] float64


def gabor_active(specs: List[GaborFilter]) -> List[GaborFilter]:
    """agabor/gabor.go:329-336."""
    return [s for s in specs if not s.Off]


def gabor_to_tensor(specs: List[GaborFilter], fs: GaborFilterSet) -> None:
    """agabor/gabor.go:89-222 ToTensor.  `for i, f := range active` copies each
    spec, so Defaults() there does not write back to the caller's slice."""
    import copy
    active = gabor_active(specs)
    nhf = 0
    nvf = 0
    if fs.Distribute:
        for f in active:
            if f.Orientation == 0:
                nhf += 1
            elif f.Orientation == 90:
                nvf += 1
    else:
        nhf = 1
        nvf = 1
    sx, sy = fs.SizeX, fs.SizeY
    radius_x = float(sx) / 2.0
    radius_y = float(sy) / 2.0
    ctr_x = float(sx - 1) / 2.0
    ctr_y = float(sy - 1) / 2.0
    h_ctr_inc = float(sy - 1) / float(nhf + 1)
    v_ctr_inc = float(sx - 1) / float(nvf + 1)
    h_cnt = 0
    v_cnt = 0
    out = np.zeros((len(active), sy, sx), dtype=np.float64)
    for i, f0 in enumerate(active):
        f = copy.copy(f0)
        f.Defaults()
        two_pi_norm = (2.0 * math.pi) / f.WaveLen
        with np.errstate(divide="ignore"):
            l_norm = float(np.float64(1.0) / np.float64(2.0 * f.SigmaLength * f.SigmaLength))
            w_norm = float(np.float64(1.0) / np.float64(2.0 * f.SigmaWidth * f.SigmaWidth))
        h_pos = 0.0
        v_pos = 0.0
        if fs.Distribute:
            if f.Orientation == 0:
                h_pos = h_ctr_inc * float(h_cnt + 1)
                h_cnt += 1
            if f.Orientation == 90:
                v_pos = v_ctr_inc * float(v_cnt + 1)
                v_cnt += 1
        else:
            h_pos = h_ctr_inc * float(h_cnt + 1)
            v_pos = v_ctr_inc * float(v_cnt + 1)
        if not f.Circular:
            for y in range(sy):
                for x in range(sx):
                    xf = float(x) - ctr_x
                    yf = float(y) - ctr_y
                    if f.Orientation == 0:
                        yf = float(y) - h_pos
                    if f.Orientation == 90:
                        xf = float(x) - v_pos
                    xfn = xf / radius_x
                    yfn = yf / radius_y
                    dist = math.hypot(xfn, yfn)
                    val = 0.0
                    if not (f.CircleEdge and dist > 1.0):
                        radians = f.Orientation * math.pi / 180
                        nx = xfn * math.cos(radians) - yfn * math.sin(radians)
                        ny = yfn * math.cos(radians) + xfn * math.sin(radians)
                        gauss = math.exp(-(w_norm * (nx * nx) + l_norm * (ny * ny)))
                        sin_val = math.sin(two_pi_norm * ny + f.PhaseOffset)
                        val = gauss * sin_val
                    out[i, y, x] = val
        else:
            norm = 1.0 / (2.0 * f.SigmaWidth * f.SigmaWidth)
            for y in range(sy):
                for x in range(sx):
                    xf = float(x) - ctr_x
                    yf = float(y) - ctr_y
                    xfn = xf / radius_x
                    yfn = yf / radius_y
                    nx = xfn * xfn * norm
                    ny = yfn * yfn * norm
                    gauss = math.sqrt(nx + ny)
                    sin_val = math.sin(two_pi_norm * nx * ny)
                    out[i, y, x] = -gauss * sin_val
    # renorm each half (gabor.go:194-221); 1/0 -> +-Inf as in Go float64
    for i in range(out.shape[0]):
        pos_sum = 0.0
        neg_sum = 0.0
        for y in range(sy):
            for x in range(sx):
                v = out[i, y, x]
                if v > 0:
                    pos_sum += v
                elif v < 0:
                    neg_sum += v
        with np.errstate(divide="ignore"):
            pos_norm = float(np.float64(1.0) / np.float64(pos_sum))
            neg_norm = float(np.float64(-1.0) / np.float64(neg_sum))
        for y in range(sy):
            for x in range(sx):
                v = out[i, y, x]
                if v > 0.0:
                    v *= pos_norm
                elif v < 0.0:
                    v *= neg_norm
                out[i, y, x] = v
    fs.Filters = out


def gabor_convolve(mel_data: np.ndarray, fs: GaborFilterSet, raw_out: np.ndarray, by_time: bool) -> bool:
    """agabor/gabor.go:225-315 Convolve.  mel_data float64 [M, S]; raw_out is a
    float32 array with 2 or 4 dims, written in place through flat stride
    arithmetic (etensor SetFloat([]int{..}) has no per-dimension bounds check).
    Returns False where the reference logs and returns without writing."""
    if mel_data.shape[1] < fs.SizeX:
        return False
    t_max = 1
    f_max = 1
    t_max_strides = 1
    nd = raw_out.ndim
    if nd == 2:
        x = mel_data.shape[1] - fs.SizeX
        if not (x == 0 or x < fs.StrideX):
            t_max = x + 1
        z = mel_data.shape[1] - fs.SizeX
        t_max_strides = z // fs.StrideX + 1
        y = mel_data.shape[0] - fs.SizeY
        if not (y == 0 or y < fs.StrideY):
            f_max = y + 1
    elif nd == 4:
        t_max = int(min(float(raw_out.shape[1] * fs.StrideX), float(mel_data.shape[1] - fs.StrideX)))
        f_max = int(min(float(raw_out.shape[0] * fs.StrideY), float(mel_data.shape[0] - fs.StrideY)))
    else:
        return False
    flat = raw_out.reshape(-1)
    strides = [int(np.prod(raw_out.shape[d + 1:])) for d in range(nd)]
    mflat = mel_data.reshape(-1)
    m_stride = mel_data.shape[1]
    nf = fs.Filters.shape[0]

    def put(idx, v):
        off = sum(i * s for i, s in zip(idx, strides))
        if off < 0 or off >= flat.size:
            raise IndexError("agabor.Convolve: output index out of range (reference panics)")
        flat[off] = np.float32(v)

    t_idx = 0
    t = 0
    while t < t_max:
        f_idx = 0
        f = 0
        while f < f_max:
            for flt in range(nf):
                f_sum = 0.0
                for ff in range(fs.SizeY):
                    for ft in range(fs.SizeX):
                        f_val = fs.Filters[flt, ff, ft]
                        moff = (f + ff) * m_stride + (t + ft)
                        if moff >= mflat.size:
                            raise IndexError("agabor.Convolve: mel index out of range (reference panics)")
                        i_val = mflat[moff]
                        if math.isnan(i_val):
                            i_val = 0.5
                        f_sum += f_val * i_val
                pos = f_sum >= 0.0
                act = fs.Gain * abs(f_sum)
                if nd == 2:
                    y = f_idx * 2
                    if by_time:
                        x = t_idx + t_max_strides * flt
                    else:
                        x = flt + t_idx * nf
                    if pos:
                        put((y, x), act)
                        put((y + 1, x), 0)
                    else:
                        put((y, x), 0)
                        put((y + 1, x), act)
                else:
                    if pos:
                        put((f_idx, t_idx, 0, flt), act)
                        put((f_idx, t_idx, 1, flt), 0)
                    else:
                        put((f_idx, t_idx, 0, flt), 0)
                        put((f_idx, t_idx, 1, flt), act)
            f += fs.StrideY
            f_idx += 1
        t += fs.StrideX
        t_idx += 1
    return True


# --------------------------------------------------------------------------
# sound.SndEnv
# --------------------------------------------------------------------------
class SndEnv:
    """sound/sndenv.go:73-182 SndEnv restricted to the speech-feature path
    (Kwta / NeighInhib are off path: SURVEY 8f).  Signal is a 1-D float64
    array; SampleRate/Channels stand in for Sound.Buf.Format."""

    def __init__(self):
        self.Params = SoundParams()
        self.DFT = DftParams()
        self.Mel = MelParams()
        self.GaborSpecs: List[GaborFilter] = []
        self.GaborFilters = GaborFilterSet()
        self.GborOutPoolsX = 0
        self.GborOutPoolsY = 0
        self.GborOutUnitsX = 0
        self.GborOutUnitsY = 0
        self.ByTime = False
        self.SampleRate = 16000
        self.Channels = 1
        self.Signal = np.zeros(0, dtype=np.float64)
        self.SegCnt = 0

    def Defaults(self):
        """sound/sndenv.go:185-192."""
        self.Params = SoundParams()
        self.Mel = MelParams()
        self.ByTime = False

    def Init(self):
        """sound/sndenv.go:195-267."""
        sr = self.SampleRate
        if sr <= 0:
            raise ValueError("sample rate <= 0")
        p = self.Params
        p.WinSamples = msec_to_samples(p.WinMs, sr)
        p.StepSamples = msec_to_samples(p.StepMs, sr)
        p.SegmentSamples = msec_to_samples(p.SegmentMs, sr)
        steps = int(go_round(p.SegmentMs / p.StepMs))
        p.SegmentSteps = steps + 2 * p.BorderSteps
        p.StrideSamples = msec_to_samples(p.StrideMs, sr)

        specs = gabor_active(self.GaborSpecs)
        nfilters = len(specs)
        gabor_to_tensor(specs, self.GaborFilters)
        if self.GborOutPoolsX == 0 and self.GborOutPoolsY == 0:
            self.GborOutput = np.zeros((self.GborOutUnitsY, self.GborOutUnitsX), dtype=np.float32)
        elif self.GborOutPoolsX > 0 and self.GborOutPoolsY > 0:
            self.GborOutput = np.zeros((self.GborOutPoolsY, self.GborOutPoolsX,
                                        self.GborOutUnitsY, self.GborOutUnitsX), dtype=np.float32)
        else:
            return
        half = p.WinSamples // 2 + 1
        self.DFT.Defaults()                                 # F7: wipes smoothing
        self.MelFilters = self.Mel.InitFilters(p.WinSamples, sr)
        self.Window = np.zeros(p.WinSamples)
        self.Power = np.zeros(half)
        self.LogPower = np.zeros(half)
        self.PowerSegment = np.zeros((half, p.SegmentSteps))
        self.LogPowerSegment = np.zeros((half, p.SegmentSteps))
        p.Steps = [p.StepSamples * (i - p.BorderSteps) for i in range(p.SegmentSteps)]
        nf = self.Mel.FBank.NFilters
        self.MelFBank = np.zeros(nf)
        self.MelFBankSegment = np.zeros((nf, p.SegmentSteps))
        self.Energy = np.zeros(p.SegmentSteps)
        if self.Mel.MFCC:
            self.MFCCSegment = np.zeros((self.Mel.NCoefs, p.SegmentSteps))
            self.MFCCDeltas = np.zeros((self.Mel.NCoefs, p.SegmentSteps))
            self.MFCCDeltaDeltas = np.zeros((self.Mel.NCoefs, p.SegmentSteps))
        siglen = len(self.Signal) - p.SegmentSamples * self.Channels
        siglen = _go_div(siglen, self.Channels)
        self.SegCnt = _go_div(siglen, p.StrideSamples) + 1

    def SndToWindow(self, start: int) -> bool:
        """sound/sndenv.go:455-478.  Returns False for the error case."""
        n = self.Params.WinSamples
        end = start + n
        if end > len(self.Signal):
            return False
        if start < 0 and end <= 0:
            self.Window = np.zeros(end - start)
        elif start < 0 and end > 0:
            self.Window = np.concatenate([np.zeros(-start), self.Signal[0:end]])
        else:
            self.Window = self.Signal[start:end]
        return True

    def ProcessStep(self, segment: int, step: int, add: int) -> bool:
        """sound/sndenv.go:438-452."""
        p = self.Params
        offset = p.Steps[step] + msec_to_samples(float(add), self.SampleRate)
        start = segment * p.StrideSamples + offset
        ok = self.SndToWindow(start)
        if ok:
            self.DFT.Filter(step, self.Window, p.WinSamples, self.Power, self.LogPower,
                            self.PowerSegment, self.LogPowerSegment)
            self.Mel.FilterDft(step, self.Power, self.MelFBankSegment, self.MelFBank, self.MelFilters)
            if self.Mel.MFCC:
                self.Mel.CepstrumDct(step, self.MelFBank, self.MFCCSegment)
        return ok

    def ProcessSegment(self, segment: int, add: int = 0):
        """sound/sndenv.go:342-433."""
        p = self.Params
        S = p.SegmentSteps
        self.Power[:] = 0
        self.LogPower[:] = 0
        self.PowerSegment[:] = 0
        self.LogPowerSegment[:] = 0
        self.Energy[:] = 0
        self.MelFBankSegment[:] = 0
        if self.Mel.MFCC:
            self.MFCCSegment[:] = 0
        for s in range(S):
            if not self.ProcessStep(segment, s, add):
                break
        # Energy (:360-366): FloatValRowCell(s, f) = Values[s*(Len/Dim0)+f], f < Dim(1)
        lp_flat = self.LogPowerSegment.reshape(-1)
        cell = lp_flat.size // self.LogPowerSegment.shape[0]
        for s in range(S):
            e = 0.0
            for f in range(self.LogPowerSegment.shape[1]):
                off = s * cell + f
                if off >= lp_flat.size:
                    raise IndexError("ProcessSegment energy: index out of range (reference panics, F6)")
                e += lp_flat[off]
            self.Energy[s] = e
        if self.Mel.MFCC:
            for s in range(S):
                self.MFCCSegment[0, s] = self.Energy[s]
        if self.Mel.MFCC and self.Mel.Deltas:
            _deltas(self.MFCCSegment, self.MFCCDeltas, self.Mel.NCoefs, S)
            _deltas(self.MFCCDeltas, self.MFCCDeltaDeltas, self.Mel.NCoefs, S)

    def ApplyGabor(self) -> np.ndarray:
        """sound/sndenv.go:481-497 with Kwta.On = NeighInhib.On = false."""
        gabor_convolve(self.MelFBankSegment, self.GaborFilters, self.GborOutput, self.ByTime)
        return self.GborOutput

    def Tail(self, signal_len: int) -> int:
        """sound/sndenv.go:503-507."""
        temp = signal_len - self.Params.SegmentSamples
        return _go_mod(temp, self.Params.StrideSamples)

    def Pad(self, signal: np.ndarray, value: float = 0.0) -> np.ndarray:
        """sound/sndenv.go:510-519."""
        tail = self.Tail(len(signal))
        pad_len = self.Params.SegmentSamples - self.Params.StepSamples - _go_mod(tail, self.Params.StepSamples)
        return np.concatenate([signal, np.full(pad_len, value, dtype=np.float64)])


def _go_div(a: int, b: int) -> int:
    """Go integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _go_mod(a: int, b: int) -> int:
    return a - b * _go_div(a, b)


def _deltas(src: np.ndarray, dst: np.ndarray, n_coefs: int, S: int):
    """sound/sndenv.go:380-404 (and :407-431): prv/nxt are reset per step s but
    carried across coefficients i; the n=2 value is the one that sticks."""
    npn = 2
    for s in range(S):
        prv = 0.0
        nxt = 0.0
        for i in range(n_coefs):
            nume = 0.0
            for n in range(1, npn + 1):
                sprv = max(s - n, 0)
                snxt = min(s + n, S - 1)
                prv += src[i, sprv]
                nxt += src[i, snxt]
                nume += float(n) * (nxt - prv)
                denom = float(2 * n * n)
                dst[i, s] = nume / denom


# --------------------------------------------------------------------------
# convenience driver used by tests / fixtures
# --------------------------------------------------------------------------
def processspeech_gabor_specs() -> List[GaborFilter]:
    """examples/processspeech/processspeech.go:226-253 parameter values."""
    specs = []
    for orient in (0.0, 45.0, 90.0, 135.0):
        for wv in (2.0,):
            for ph in (0.0, 1.5708):
                for sg in (0.5,):
                    specs.append(GaborFilter(WaveLen=wv, Orientation=orient, SigmaWidth=sg,
                                             SigmaLength=sg, PhaseOffset=ph, CircleEdge=True))
    return specs


def make_env(signal: np.ndarray, sample_rate: int = 16000, mfcc: bool = False, deltas: bool = False,
             gabor: bool = True, prev_smooth: float = 0.0, cur_smooth: Optional[float] = None,
             out4d: bool = True, by_time: bool = False) -> SndEnv:
    """SndEnv configured like BASELINE config 1 (processspeech gabor set,
    4-D out [8,2,2,8]); smoothing is applied after Init (F7)."""
    se = SndEnv()
    se.Defaults()
    se.SampleRate = sample_rate
    se.Signal = np.asarray(signal, dtype=np.float64)
    se.Mel.MFCC = mfcc
    se.Mel.Deltas = deltas
    if gabor:
        se.GaborSpecs = processspeech_gabor_specs()
        se.GaborFilters.SizeX = 9
        se.GaborFilters.SizeY = 9
        se.GaborFilters.StrideX = 3
        se.GaborFilters.StrideY = 3
        se.GaborFilters.Gain = 2
        se.GaborFilters.Distribute = False
        se.ByTime = by_time
        if out4d:
            se.GborOutPoolsY, se.GborOutPoolsX, se.GborOutUnitsY, se.GborOutUnitsX = 8, 2, 2, 8
        else:
            se.GborOutPoolsX = se.GborOutPoolsY = 0
            se.GborOutUnitsY, se.GborOutUnitsX = 16, 16
    else:
        se.GaborFilters.SizeX = se.GaborFilters.SizeY = 1
        se.GaborFilters.StrideX = se.GaborFilters.StrideY = 1
        se.GborOutUnitsX = se.GborOutUnitsY = 1
    se.Init()
    se.DFT.PrevSmooth = prev_smooth
    se.DFT.CurSmooth = (1.0 - prev_smooth) if cur_smooth is None else cur_smooth
    return se


def process_all(se: SndEnv, add: int = 0, want_gabor: bool = True, want_power: bool = False):
    """Run every segment; returns dict of stacked per-segment outputs."""
    out = {"mel": [], "energy": []}
    if se.Mel.MFCC:
        out["mfcc"] = []
        if se.Mel.Deltas:
            out["deltas"] = []
            out["delta_deltas"] = []
    if want_gabor:
        out["gabor"] = []
    if want_power:
        out["power"] = []
        out["logpower"] = []
    for seg in range(se.SegCnt):
        se.ProcessSegment(seg, add)
        out["mel"].append(se.MelFBankSegment.copy())
        out["energy"].append(se.Energy.copy())
        if se.Mel.MFCC:
            out["mfcc"].append(se.MFCCSegment.copy())
            if se.Mel.Deltas:
                out["deltas"].append(se.MFCCDeltas.copy())
                out["delta_deltas"].append(se.MFCCDeltaDeltas.copy())
        if want_gabor:
            out["gabor"].append(se.ApplyGabor().copy())
        if want_power:
            out["power"].append(se.PowerSegment.copy())
            out["logpower"].append(se.LogPowerSegment.copy())
    return {k: np.stack(v) if len(v) else np.zeros((0,)) for k, v in out.items()}
