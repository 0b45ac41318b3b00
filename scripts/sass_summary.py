#!/usr/bin/env python3
"""Regenerates profiles/r02_sass_summary.md: per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths
(cuobjdump -sass on the in-tree libtb_b200.so).  python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'trackingbench_slam_b200', 'libtb_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', txt)), capture_output=True, text=True).stdout.split('\n')
cols = ['UTCIMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'ELECT', 'VABSDIFF4', 'VIMNMX3', 'IDP.4A', 'IDP.2A']
rows = []
for i, blk in enumerate(txt.split('Function : ')[1:]):
    ins = re.findall(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', blk, flags=re.M)
    name = names[i] if i < len(names) else blk.split('\n')[0]
    name = re.sub(r'\((bool|int)\)', '', name)
    name = re.sub(r'\(.*', '', name)
    rows.append((len(ins), name, [sum(1 for x in ins if x.startswith(c)) for c in cols]))
rows.sort(reverse=True)
print('# SASS evidence (cuobjdump -sass trackingbench_slam_b200/libtb_b200.so, sm_100a) — instruction counts per kernel\n')
print('Mnemonics: `UTCIMMA` = tcgen05.mma (int8), `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = cp.async.bulk.tensor (TMA tiled load), `UBLKCP` = cp.async.bulk,')
print('`LDGSTS` = cp.async, `SYNCS` = mbarrier ops, `ELECT` = elect.sync, `VABSDIFF4` / `VIMNMX3` = byte-SIMD absolute difference / three-input min-max,')
print('`IDP.4A` / `IDP.2A` = DP4A / DP2A.  Regenerate with `python scripts/sass_summary.py`.\n')
print('| kernel | SASS instructions | ' + ' | '.join(cols) + ' |')
print('|---|---|' + '---|' * len(cols))
for n, name, c in rows:
    if sum(c) == 0: continue
    print('| `%s` | %d | %s |' % (name, n, ' | '.join(str(x) for x in c)))
