#!/usr/bin/env python
"""Search lane mappings / strides of the fused kernel's shared-memory accesses for bank conflicts."""
import itertools, sys
sys.path.insert(0, 'tools')
from bankconf import wavefronts

WIN = 264

def lane_map(idle):
    m, k = [], 0
    for lane in range(32):
        if lane in idle: m.append(None)
        else:
            m.append((k // 10, k % 10)); k += 1
    return m

def cost(idle, RS, PS, twp, verbose=False):
    L = lane_map(idle)
    def w(fn, width):
        return wavefronts([None if l is None else fn(*l) for l in L], width)
    c = {}
    c['win'] = 40 * w(lambda q, j: 8 * (q * PS + WIN + j), 8)                      # ideal 2
    c['tw'] = 19 * w(lambda q, j: 16 * j if twp == 10 else 16 * (j + (j // 8) * (twp - 10)), 16)  # broadcast across q
    c['est'] = 20 * w(lambda q, j: 8 * (q * PS + 2 * j), 16)                        # ideal 4
    c['ldA'] = 10 * w(lambda q, j: 8 * (q * PS + RS * j), 16)
    c['ldB'] = 10 * w(lambda q, j: 8 * (q * PS + RS * (10 if j == 0 else 20 - j)), 16)
    def pw(fn):
        return wavefronts([None if (l is None or l[1] == 0) else fn(*l) for l in L], 8)
    c['p1'] = 10 * pw(lambda q, j: 8 * (q * PS + j))
    c['p2'] = 10 * pw(lambda q, j: 8 * (q * PS + 20 - j))
    tot = sum(c.values())
    if verbose: print(c)
    return tot

if __name__ == "__main__":
    res = []
    for idle in itertools.combinations(range(32), 2):
        for RS in (22, 26):
            for PS in range(544, 600, 2):
                res.append((cost(idle, RS, PS, 10), idle, RS, PS))
    res.sort()
    for r in res[:10]: print(r)
    print("current:", cost((30, 31), 22, 554, 10, True))
    print("best   :", cost(res[0][1], res[0][2], res[0][3], 10, True))
