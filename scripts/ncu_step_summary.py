#!/usr/bin/env python3
"""Turns the ncu launch list of one bench step (csv from
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline)
into profiles/r02_ncu_step_summary.json (per-kernel launches, time, share, DRAM bytes of the LAST step in the list) and
profiles/r02_traffic.json (DRAM bytes per kernel and step, what bench.py reports as roofline.traffic).
usage: python scripts/ncu_step_summary.py launches.csv"""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit')}
launches = {}
for r in rows[1:]:
    if len(r) < len(hdr): continue
    d = launches.setdefault(int(r[ix['ID']]), {'name': re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').strip()})
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    if 'time_duration' in r[ix['Metric Name']]: d['us'] = v / 1e3 if u in ('ns', 'nsecond') else v * (1e3 if u.startswith('ms') else 1)
    if 'bytes_read' in r[ix['Metric Name']]: d['rd'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    if 'bytes_write' in r[ix['Metric Name']]: d['wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
ids = sorted(launches)
# the last step = everything from the last launch of the first pyramid level's kernel chain: find the last run of launches that starts with a pyr kernel after a matcher kernel
names = [launches[i]['name'] for i in ids]
start = 0
for j in range(1, len(ids)):
    if names[j].startswith('pyr_resize') and not names[j - 1].startswith('pyr_resize'): start = j
step = [launches[i] for i in ids[start:]]
per = {}
for d in step:
    k = re.sub(r'<.*', '', d['name'])
    p = per.setdefault(k, {'launches': 0, 'us': 0.0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0})
    p['launches'] += 1; p['us'] += d.get('us', 0); p['dram_read_bytes'] += d.get('rd', 0); p['dram_write_bytes'] += d.get('wr', 0)
tot = sum(p['us'] for p in per.values())
for p in per.values(): p['share'] = p['us'] / tot
json.dump({'what': 'one 256-frame 1080p step under ncu (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none): cold-cache, serialised launches; shares agree with the CUDA-event table, absolutes do not',
           'per_kernel': per, 'total_us': tot}, open(os.path.join(ROOT, 'profiles', 'r02_ncu_step_summary.json'), 'w'), indent=1)
alias = {'pyr_resize_linear_u8_stream8': 'pyr_resize', 'pyr_resize_linear_u8_stream': 'pyr_resize', 'pyr_resize_linear_u8_tma': 'pyr_resize'}
tr = {}
for k, p in per.items():
    kk = alias.get(k, k)
    tr[kk] = tr.get(kk, 0) + p['dram_read_bytes'] + p['dram_write_bytes']
tr['_note'] = 'dram__bytes_read.sum + dram__bytes_write.sum per 256-frame step (all launches of the kernel in one step), ncu --clock-control none, profiles/r02_ncu_launches.csv'
json.dump(tr, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)
print(json.dumps({k: (round(p['us'], 1), p['launches'], round(p['share'], 3)) for k, p in per.items()}))
