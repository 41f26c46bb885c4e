"""Attributes warp stall samples (time) to CUDA source lines: ncu_time.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}; tot = 0; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] == '': continue
    i = hdr.index('# Samples')
    v = int(r[i]) if r[i].isdigit() else 0
    key = (cur, int(r[0]), r[1].strip()[:100]); per[key] = per.get(key, 0) + v; tot += v
print("total samples", tot)
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    print("%5.1f%%  %s:%d  %s" % (100.0 * v / tot, k[0], k[1], k[2]))
if len(sys.argv) > 3:
    # line-range categories: name:lo-hi,...
    cats = []
    for item in sys.argv[3].split(','):
        name, rng = item.split(':'); lo, hi = rng.split('-'); cats.append((name, int(lo), int(hi)))
    agg = {}
    for (f, ln, _), v in per.items():
        name = 'other:' + f
        if f == 'sampler_spec.cu':
            name = 'spec:other'
            for nm, lo, hi in cats:
                if lo <= ln <= hi: name = nm; break
        agg[name] = agg.get(name, 0) + v
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print("%5.1f%%  %s" % (100.0 * v / tot, k))
