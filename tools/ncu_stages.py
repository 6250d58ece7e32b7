#!/usr/bin/env python
"""Per-stage instruction / sample / stall summary of an ncu source-page CSV for the fused kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2] if len(sys.argv) > 2 else 'auditory_b200/csrc/aud_kernels.cuh'
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]; ix = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
wf = hdr.index('L1 Wavefronts Shared'); wfi = hdr.index('L1 Wavefronts Shared Ideal')
stalls = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
sidx = {c: hdr.index(c) for c in stalls}
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        lines.append((int(r[0]), int(r[ix]), int(r[isamp]), {c: int(r[sidx[c]]) for c in stalls}, int(r[wf]), int(r[wfi])))
    except ValueError:
        pass
tot = sum(l[1] for l in lines); ts = sum(l[2] for l in lines)
src = open(srcfile).read().split('\n')
def find(s):
    return [i + 1 for i, l in enumerate(src) if s in l][0]
names = [("dft helpers", '__device__ __forceinline__ void dft4'), ("other helpers", '// ------------------------------------------------------------ small helpers'),
         ("carve etc", 'struct Smem {'), ("records", '// ------------------------------------------------------------ frame-pair records'),
         ("fft: stage", 'auto resolve_own = '), ("fft: round head+wait", 'PairInfo own_cur, own_nxt;'),
         ("pass1 loads", 'float ar[20], ai[20], br[20], bi[20];'), ("pass1 dft+tw+st", '// pass 1: columns n2'),
         ("pass2 loads", '// pass 2: a lane transforms'), ("prefetch call", 'if constexpr (EPIREC) stage(R + 1, rnext);'), ("pass2 dft+pairing", '// |X_A|^2, |X_B|^2 of bin k'),
         ("coop selfpair", '// ---- the self-paired columns'), ("empty wait+rlow", '// the ring slots of this round were last used'), ("mel", '// ---- mel filter bank on the raw'), ("fft: round tail", 'if (lane == 0) mbar_arrive(&full[R & 1]);   // release'),
         ("tile stage", '// gabor weights [nf][sy][sx] (agabor.ToTensor layout)'),
         ("epilogue", '// ------------------------------------------------------------ epilogue warps'), ("kernel main", '// ------------------------------------------------------------ fused kernel'), ("end", '// ------------------------------------------------- power / log-power')]
marks = [(n, find(t)) for n, t in names]
print(f"total warp-inst {tot/1e6:.1f}M  samples {ts}")
print(f"{'stage':20s} {'inst%':>6s} {'Minst':>7s} {'samp%':>6s} {'wf(M)':>7s} {'ideal':>7s}  top stalls")
for k in range(len(marks) - 1):
    a, b = marks[k][1], marks[k + 1][1]
    sel = [l for l in lines if a <= l[0] < b]
    i = sum(l[1] for l in sel); s = sum(l[2] for l in sel); w = sum(l[4] for l in sel); wi = sum(l[5] for l in sel)
    st = {c: sum(l[3][c] for l in sel) for c in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{marks[k][0]:20s} {100*i/tot:6.1f} {i/1e6:7.1f} {100*s/ts:6.1f} {w/1e6:7.2f} {wi/1e6:7.2f}  " + ", ".join(f"{kk[6:]}={100*v/max(s,1):.0f}%" for kk, v in top))
