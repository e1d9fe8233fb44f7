#!/usr/bin/env python
"""Correlates an `ncu --page source --csv` SASS export with source lines via `nvdisasm -g -c`.
usage: tools_sass_lines.py <src.csv> <dis.txt> <mangled kernel name> [top]"""
import csv, re, sys, collections
src, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
SORTK = 0 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 1
rows = list(csv.reader(open(src)))
hdr = rows[1]
ie, isamp, ith = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
insts = [(r[1].strip(), float(r[ie] or 0), float(r[isamp] or 0), float(r[ith] or 0)) for r in rows[2:] if len(r) > ie]
# line info from nvdisasm
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':'))
cur = ('?', 0)
seq = []
for l in lines[start + 1:]:
    if l.startswith('//-----') or l.startswith('.text.'):
        if seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        # inlined-at chains: keep the innermost (first) record
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append(cur)
assert len(seq) >= len(insts), (len(seq), len(insts))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (txt, n, s, th), loc in zip(insts, seq):
    a = agg[loc]; a[0] += n; a[1] += s; a[2] += th
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print('total warp-inst %.3e  samples %d' % (tot, tots))
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][SORTK])[:top]:
    if f not in srcs:
        try: srcs[f] = open('/root/repo/icm_slam_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ''
    print('%5.1f%% inst %5.1f%% stall-samples  thr/inst %4.1f  %s:%d  %s' % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / max(a[0], 1), f, ln, text))
