"""The shipped path must not route through the oracle or any CPU fallback: nothing under auditory_b200/,
include/ or go/ may import, link or call oracle/ (only tests/, bench.py's CPU legs and
__graft_entry__.smoke() use it as the checker), and the shared library must not contain oracle symbols."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sources(*dirs, exts=(".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".go", ".txt")):
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            if "__pycache__" in base or os.sep + "lib" in base:
                continue
            for f in files:
                if f.endswith(exts) or f == "Makefile":
                    yield os.path.join(base, f)


def test_product_sources_never_touch_the_oracle():
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|liboracle|orc_[a-z_]+\s*\(|np_oracle|c_oracle)")
    bad = []
    for path in _sources("auditory_b200", "include", "go"):
        for n, line in enumerate(open(path, errors="replace"), 1):
            if pat.search(line):
                bad.append(f"{os.path.relpath(path, ROOT)}:{n}: {line.strip()}")
    assert not bad, "\n".join(bad)


def test_shared_library_has_no_oracle_symbols_and_no_cpu_compute_entry():
    lib = os.path.join(ROOT, "auditory_b200", "lib", "libauditory_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert not [s for s in syms if s.startswith("orc_")]
    exported = sorted(s for s in syms if s.startswith("aud_"))
    from auditory_b200 import _lib
    assert exported == sorted(_lib.SYMBOLS), (exported, sorted(_lib.SYMBOLS))


def test_bench_product_arm_only_uses_oracle_for_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    # the oracle is imported inside the CPU-baseline / reference-arm helpers only
    for m in re.finditer(r"^(\s*)from oracle import", src, flags=re.M):
        assert len(m.group(1)) >= 4, "bench.py imports the oracle at module level"
