"""Dumps the SASS of a kernel from a .ncu-rep with per-instruction executed count (per warp-step),
stall samples and the dominant stall reason: ncu_sass.py rep warps steps [min_exec_per_step]"""
import csv, subprocess, sys
rep, warps, steps = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    ex = float(r[ix['Instructions Executed']] or 0) / (warps * steps)
    smp = int(r[ix['# Samples']] or 0)
    tot_s += smp
    if ex < thr: continue
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    bc = r[ix['L1 Wavefronts Shared Excessive']]
    print("%5.2f %6d %-22s %s%s" % (ex, smp, ','.join('%s:%d' % (n, v) for v, n in top if v), r[ix['Source']].strip()[:100],
                                   ('   [smem excess wf %s]' % bc) if bc not in ('0', '-', '') else ''))
print("total samples", tot_s)
