#!/usr/bin/env python
"""Shared-memory bank-conflict model used to pick the exchange-buffer layout.
32 banks x 4 B; a request of w bytes/lane is served in groups of 128/w lanes;
wavefronts per group = max over banks of distinct 4-byte words hitting it."""
import itertools

def wavefronts(addrs_bytes, width):
    """addrs_bytes: per-lane byte address or None (inactive); width in {4,8,16}."""
    lanes_per = 128 // width
    tot = 0
    for g in range(0, 32, lanes_per):
        words = {}
        for a in addrs_bytes[g:g + lanes_per]:
            if a is None: continue
            for w in range(a // 4, a // 4 + width // 4):
                words.setdefault(w % 32, set()).add(w)
        tot += max((len(s) for s in words.values()), default=0)
    return tot

def lanes():
    for lane in range(32):
        yield (lane // 10, lane % 10) if lane < 30 else None

def check(RS, PS, verbose=False):
    """k2-fastest layout: slot(k1, x) = k1*RS + x (float2 units), pair stride PS."""
    res = {}
    # B: pass-1 store, STS.128 of columns n2=2j,2j+1 at row k1 (any k1: constant offset)
    a = [None if l is None else 8 * (l[0] * PS + 2 * l[1]) for l in lanes()]
    res['st128'] = wavefronts(a, 16)
    # C: pass-2 load row k1 = j + 10c, LDS.128 at n2 (const)
    a = [None if l is None else 8 * (l[0] * PS + l[1] * RS) for l in lanes()]
    res['ld128'] = wavefronts(a, 16)
    return res

if __name__ == "__main__":
    best = []
    for RS in range(20, 33, 2):
        for PS in range(20 * RS, 20 * RS + 33, 2):
            r = check(RS, PS)
            best.append((r['st128'] + r['ld128'], RS, PS, r))
    best.sort(key=lambda t: (t[0], t[2]))
    for b in best[:12]:
        print(b)
