#!/usr/bin/env python
"""Stall-sample summary of an ncu report's SASS page: tools/sass_stalls.py rep.ncu-rep [top_n]"""
import csv, re, collections, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) == len(hdr): data.append(r)
S = lambda r, k: int(r[ix[k]] or 0)
tot = sum(S(r, '# Samples') for r in data)
print('total samples', tot, 'inst', sum(S(r, 'Instructions Executed') for r in data))
for k in ['stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_selected', 'stall_not_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_sleep', 'stall_lg', 'stall_math', 'stall_mio', 'stall_dispatch']:
    v = sum(S(r, k) for r in data); print(f"{k:24s}{v:7d} {100*v/max(tot,1):5.1f}%")
for r in sorted(data, key=lambda r: -S(r, '# Samples'))[:topn]:
    print(S(r, '# Samples'), S(r, 'Instructions Executed'), r[0][-5:], r[1][:64], 'bar', S(r, 'stall_barrier'), 'lsb', S(r, 'stall_long_sb'), 'ssb', S(r, 'stall_short_sb'), 'wait', S(r, 'stall_wait'), 'sleep', S(r, 'stall_sleep'))
blk = collections.OrderedDict()
for i, r in enumerate(data):
    d = blk.setdefault(i // 64, [r[0][-5:], 0, 0, set()])
    d[1] += S(r, '# Samples'); d[2] = max(d[2], S(r, 'Instructions Executed'))
    op = re.sub(r'@!?U?P\d\s+', '', r[1]).split()[0].split('.')[0]
    if op in ('VOTE', 'ATOMG', 'LDG', 'BAR', 'ATOMS', 'NANOSLEEP', 'B2R', 'LDL', 'STL'): d[3].add(op)
for b, d in blk.items():
    if d[1] > tot / 150:
        print(d[0], 'samples', d[1], 'maxexec', d[2], sorted(d[3]))
