#!/usr/bin/env python
"""Print the measured FP32 (non-tensor) peaks of cuda:0 (aud_measure_fp32) as one JSON object."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auditory_b200 import _lib

def measure(device=0):
    L = _lib.lib()
    names = {0: "ffma", 1: "ffma2_packed", 2: "mix_fadd_fmul_ffma_2_1_1", 3: "mix_packed"}
    out = {}
    for kind, name in names.items():
        t, g = C.c_double(0), C.c_double(0)
        _lib.check(L.aud_measure_fp32(device, kind, C.byref(t), C.byref(g)))
        out[name] = {"tflops": t.value, "lane_ginst_per_s": g.value}
    return out

if __name__ == "__main__":
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 0)))
