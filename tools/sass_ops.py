"""Dynamic opcode histogram of a gzip'd `ncu --page source --csv --print-source sass` export: python tools/sass_ops.py file.csv.gz [units]
(units = what to divide the executed counts by, e.g. the number of 64-env chunks of the launch)."""
import csv, gzip, collections, re, sys
rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ci = {n: i for i, n in enumerate(rows[hi])}
ops = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    try:
        ie = int(r[ci["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
    op = m.group(2) if m else "?"
    base = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "I2F", "F2I", "F2F", "IMAD", "UTC", "LDTM", "STTM", "SYNCS")) else op.split(".")[0]
    ops[base] += ie; tot += ie
print(rows[0][1][:120] if rows[0] else "")
print("total warp-instructions %d  (%.1f per unit)" % (tot, tot / units))
for k, v in ops.most_common(50):
    print("%-16s %11d  %9.1f per unit  %5.1f%%" % (k, v, v / units, 100.0 * v / tot))
