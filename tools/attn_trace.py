#!/usr/bin/env python
"""Timeline of one CTA of the head-resident attention kernels (PEVIT_ATTN_TRACE=<prefix> python tools/attn_bench.py ...
writes <prefix>.fwd.bin / <prefix>.bwd.bin; each is overwritten by the latest launch).
usage: attn_trace.py file.bin [first_event last_event]   -- per warp: mean cycles of every (event -> next event)
transition, then the merged timeline of events [first, last) of the MMA warp's numbering."""
import struct
import sys
from collections import defaultdict

NAMES = {1: "mma:operands landed", 2: "mma:S/MMA1 issued", 3: "mma:P ready (p_full/pds_full)", 4: "mma:PV/MMA2 issued",
         5: "mma:acc drained (o_empty/kv_empty)", 6: "mma:dq drained", 9: "wg:head operands landed",
         10: "wg:S ready", 11: "wg:loaded+max", 12: "wg:after named barrier", 13: "wg:exp done, st issued",
         14: "wg:arrived", 15: "wg:ds_free", 20: "out:acc full", 21: "out:kv drained", 22: "out:stored",
         23: "out:dq full", 24: "out:dq drained", 30: "tma:stage free, loads issued", 40: "mma:S mmas issued", 41: "mma:S committed",
         42: "mma:fence done", 43: "mma:PV/dV mmas issued", 44: "mma:dK mmas issued", 45: "mma:dQ mmas issued", 16: "wg:delta computed", 17: "wg:delta barrier passed",
         50: "kernel entry", 51: "prologue done", 52: "role done", 21: "out:accumulators released",
         46: "mma:first MMA1 of head issued", 47: "mma:second MMA1 of head issued"}


def main():
    raw = open(sys.argv[1], "rb").read()
    warps = struct.unpack_from("<i", raw, 0)[0]
    n = (len(raw) - 4) // 8 // warps
    ev = []
    for w in range(warps):
        vals = struct.unpack_from(f"<{n}Q", raw, 4 + 8 * n * w)
        ev.append([(v >> 8, v & 0xff) for v in vals if v])
    t0 = min(e[0][0] for e in ev if e)
    for w, e in enumerate(ev):
        if not e:
            continue
        trans = defaultdict(list)
        for (ta, a), (tb, b) in zip(e, e[1:]):
            trans[(a, b)].append(tb - ta)
        print(f"--- warp {w}: {len(e)} events, span {e[-1][0] - e[0][0]} cycles")
        for (a, b), d in sorted(trans.items(), key=lambda kv: -sum(kv[1])):
            d2 = sorted(d)
            print(f"   {a:>2}->{b:<2} n={len(d):4d} mean={sum(d) / len(d):8.0f} med={d2[len(d2) // 2]:7d} max={d2[-1]:7d} total={sum(d):9d}   "
                  f"{NAMES.get(a, '?')} -> {NAMES.get(b, '?')}")
    if len(sys.argv) > 3:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        merged = sorted((t - t0, w, i) for w, e in enumerate(ev) for (t, i) in e)
        for t, w, i in merged:
            if lo <= t < hi:
                print(f"{t:9d}  w{w:<2d} {i:>2} {NAMES.get(i, '?')}")


if __name__ == "__main__":
    main()
