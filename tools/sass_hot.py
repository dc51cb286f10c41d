"""Read a gzip'd `ncu --page source --csv --print-source sass` export: top stall-sample instructions and totals by stall reason."""
import csv, gzip, sys, collections
rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
data = rows[hi + 1:]
ci = {n: i for i, n in enumerate(h)}
S, NS, IE, SRC = ci["# Samples"], ci["Warp Stall Sampling (Not-issued Samples)"], ci["Instructions Executed"], ci["Source"]
stall_cols = [n for n in h if n.startswith("stall_") and "(Not Issued)" not in n]
tot = collections.Counter()
tsamp = 0
recs = []
for k, r in enumerate(data):
    try:
        s = int(r[S])
    except (ValueError, IndexError):
        continue
    tsamp += s
    d = {n: int(r[ci[n]] or 0) for n in stall_cols}
    for n, v in d.items():
        tot[n] += v
    recs.append((s, k, r[SRC].strip(), int(r[IE] or 0), d))
print("total samples", tsamp)
print("by reason:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / tsamp) for n, v in tot.most_common(10)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("top instructions by samples:")
for s, k, src, ie, d in sorted(recs, reverse=True)[:top]:
    why = ", ".join("%s %d" % (n[6:], v) for n, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v)
    print("%6d (%4.1f%%)  #%5d  exec %8d  %-60s %s" % (s, 100.0 * s / tsamp, k, ie, src[:60], why))
if len(sys.argv) > 3:          # dump a window of instructions: start end
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for s, k, src, ie, d in recs:
        if a <= k <= b:
            print("#%5d %6d exec %8d  %s" % (k, s, ie, src))
