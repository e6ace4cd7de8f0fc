"""Top stall-sample SASS instructions per kernel of an .ncu-rep captured with --import-source on:
python tools/ncu_top_stalls.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern, cur, seen = [], None, set()
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        kern.append(cur)
        continue
    if cur is None:
        continue
    if r and r[0] == 'Address':
        cur['hdr'] = r
        continue
    cur['rows'].append(r)
for k in kern:
    if k['name'] in seen:
        continue
    seen.add(k['name'])
    h = k['hdr']
    ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    tot = sum(int(r[isamp] or 0) for r in k['rows'] if len(r) > isamp)
    print(k['name'][:100], '| samples', tot, '| SASS instructions', len(k['rows']))
    top = sorted([(int(r[isamp] or 0), i, r[ia], r[iex]) for i, r in enumerate(k['rows']) if len(r) > isamp], reverse=True)[:top_n]
    for s, i, src, ex in top:
        print("   %6d %5.1f%% sass#%4d executed=%s  %s" % (s, 100.0 * s / max(tot, 1), i, ex, src[:110]))
