"""Summarise a DGP_TRACE file: durations of the P-chain pieces and the T launches (development aid)."""
import sys, collections
rows = [l.strip().split(",") for l in open(sys.argv[1]) if not l.startswith("#")]
ev = [(int(a), b, int(c), float(d)) for a, b, c, d in rows]
last = {0: None, 1: None}
dur = collections.defaultdict(list)
for lane, tag, arg, t in ev:
    if last[lane] is not None and tag.endswith(">"):
        dur[(lane, tag)].append((arg, t - last[lane]))
    last[lane] = t
for k, v in sorted(dur.items()):
    d = [x[1] for x in v]
    d2 = sorted(d)
    print(k, "count", len(d), "sum %.1f us" % sum(d), "median %.1f" % d2[len(d2) // 2], "min %.1f max %.1f" % (d2[0], d2[-1]))
print("end of trace: %.1f us" % max(e[3] for e in ev))
if len(sys.argv) > 2:
    for k, v in sorted(dur.items()):
        print(k, " ".join("%d:%.0f" % x for x in v))
