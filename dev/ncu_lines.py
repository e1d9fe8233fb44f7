"""per-source-line instruction / stall-sample totals from `ncu --page source --csv --print-source cuda,sass` output"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
acc = collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(k):
        try: return int(r[hdr[k]])
        except Exception: return 0
    key = (cur_file, ln)
    a = acc.setdefault(key, [0, 0, r[1][:110], 0, 0, 0])
    a[0] += num("Instructions Executed"); a[1] += num("# Samples")
    for j, k in ((3, "stall_long_sb"), (4, "stall_barrier"), (5, "stall_short_sb")):
        if k in hdr: a[j] += num(k)
ti = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
print("total warp instr", ti, "samples", ts)
print("  instr%  smp%  long% barr% short%  file:line  source")
for (f, ln), a in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.1f %6.1f %5.1f %5.1f %5.1f  %s:%d  %s" % (100 * a[0] / ti, 100 * a[1] / ts, 100 * a[3] / ts, 100 * a[4] / ts, 100 * a[5] / ts, f, ln, a[2].strip()))
