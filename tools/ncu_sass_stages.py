#!/usr/bin/env python
"""Per-stage dynamic summary of a fused-kernel ncu capture from its SASS page, with each SASS address mapped to a
source (file, line) through `nvdisasm -g` of the same kernel built from the same sources.

  ncu -i rep.ncu-rep --page source --print-source sass --csv > sass.csv
  python tools/ncu_sass_stages.py sass.csv NW NE ER        # shape of the captured kernel, e.g. 12 4 1
"""
import bisect, collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "auditory_b200", "csrc")
sass_csv, nw, ne, er = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
tmp = tempfile.mkdtemp()
cubin = os.path.join(tmp, "k.cubin")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                       "-I", os.path.join(ROOT, "include"), "-I", CS, f"-DAUD_NW={nw}", f"-DAUD_NE={ne}", f"-DAUD_ER={er}",
                       "-cubin", "-o", cubin, os.path.join(CS, "aud_fused_variant.cu")])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# address -> innermost (file, line) and the outermost aud_kernels.cuh line of the inline chain
addr_line = {}
cur, outer = None, None
infn = False
for l in dis.split("\n"):
    if l.startswith(".text."):
        infn = "fused_features_kernel" in l
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        if m.group(3):
            outer = (os.path.basename(m.group(3)), int(m.group(4)))
        else:
            outer = cur
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and infn:
        addr_line[int(m.group(1), 16)] = (cur, outer, m.group(2))

rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {c: hdr.index(c) for c in hdr}
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
base = None
recs = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16)
    if base is None:
        base = a
    recs.append((a - base, r))

def stage_of(loc):
    (f, ln), (of, oln), txt = loc
    if f == "aud_fft_core.cuh":
        return "core:" + core_fn(ln)
    if f == "aud_kernels.cuh":
        return "k:" + kern_fn(ln)
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    return f"hdr:{f}:{op.split('.')[0]}"     # CUDA header intrinsics (no inline chain in the line table)

def marks_of(path, pats):
    src = open(path).read().split("\n")
    out = []
    for name, pat in pats:
        idx = [i + 1 for i, l in enumerate(src) if pat in l]
        if idx:
            out.append((idx[0], name))
    return sorted(out)

core_marks = marks_of(os.path.join(CS, "aud_fft_core.cuh"), [
    ("packed ops", "AUD_HD f2 add2"), ("dft20", "AUD_HD void dft4"), ("twiddle consts", "AUD_HD float w400r"),
    ("layout/assign", "// ------------------------------------------------------------------ exchange layout"),
    ("twiddle_row", "AUD_HD void twiddle_row"), ("pass1_store", "AUD_HD void pass1_store"), ("pass2_load", "AUD_HD void pass2_load"),
    ("pass2_power", "AUD_HD void pass2_power"), ("pass2_park", "AUD_HD void pass2_park"), ("selfpair", "AUD_HD void selfpair_item"),
    ("frame_peak", "AUD_HD float max3_nan")])
kern_marks = marks_of(os.path.join(CS, "aud_kernels.cuh"), [
    ("helpers", "// ------------------------------------------------------------ small helpers"), ("carve", "struct Smem {"),
    ("finish_mel/ring_slot", "__device__ __forceinline__ int ring_slot"),
    ("records", "// ------------------------------------------------------------ frame-pair records"),
    ("fft: prologue/stage", "__device__ __forceinline__ void fft_role"), ("fft: round head", "for (int R = 0; R < rounds; ++R) {\n"),
    ("fft: window loads", "f2 xr[20], xi[20];"), ("fft: levels", "// ---- frame levels."), ("fft: alone reload", "if (rep == 1) {"),
    ("fft: pass1 call", "// pass 1: columns 2j"), ("fft: pass2 call", "// pass 2: the lane owns row pair"),
    ("fft: selfpair call", "// ---- the self-paired rows 0 and 10"), ("fft: empty wait", "// the ring slots of this round were last used"),
    ("fft: rlow/rawpow", "// ---- low bins for Energy"), ("fft: mel", "// ---- mel filter bank on the raw"),
    ("fft: flags/tail", "// the spare column of a frame's ring row"), ("tile stage", "// gabor weights [nf][sy][sx]"),
    ("epilogue", "// ------------------------------------------------------------ epilogue warps"),
    ("kernel main", "// ------------------------------------------------------------ fused kernel")])
def pick(marks, ln):
    i = bisect.bisect_right([m[0] for m in marks], ln) - 1
    return marks[i][1] if i >= 0 else "head"
core_fn = lambda ln: pick(core_marks, ln)
kern_fn = lambda ln: pick(kern_marks, ln)

agg = collections.defaultdict(lambda: collections.Counter())
unk = 0
for off, r in recs:
    loc = addr_line.get(off)
    if loc is None or loc[0] is None:
        unk += int(r[col["Instructions Executed"]] or 0)
        st = "unknown"
    else:
        st = stage_of(loc)
    a = agg[st]
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    a["samp"] += int(r[col["# Samples"]] or 0)
    a["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    a["wfi"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    for s in stalls:
        a[s] += int(r[col[s]] or 0)
tot = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print(f"total warp-inst {tot/1e6:.1f}M  samples {ts}  (unmapped {unk/1e6:.2f}M)")
print(f"{'stage':26s} {'inst%':>6s} {'Minst':>7s} {'samp%':>6s} {'wf(M)':>7s} {'ideal':>7s}  top stalls")
for st, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
    top = sorted(((s, a[s]) for s in stalls), key=lambda kv: -kv[1])[:4]
    print(f"{st:26s} {100*a['inst']/tot:6.1f} {a['inst']/1e6:7.1f} {100*a['samp']/max(ts,1):6.1f} {a['wf']/1e6:7.2f} {a['wfi']/1e6:7.2f}  "
          + ", ".join(f"{k[6:]}={100*v/max(a['samp'],1):.0f}%" for k, v in top))
