"""Per source line of one file: executed warp instructions and stall samples (needs -lineinfo): ncu_stalls.py rep file.cu [top]"""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; per = {}
for r in rows:
    if len(r) == 2 and r[0] in ('File Path', 'File Name'): cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or cur != fname or r[0] == '': continue
    def col(name):
        v = r[hdr.index(name)]
        return int(v) if v.isdigit() else 0
    stalls = {h: col(h) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    per[int(r[0])] = (r[1].strip()[:100], col('# Samples'), col('Instructions Executed'), stalls)
tot = sum(v[1] for v in per.values()) or 1
print("total samples %d" % tot)
for ln, (src, n, ins, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    big = ' '.join('%s=%d' % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print("%5.1f%% %8d instr  L%-4d %-100s  %s" % (100.0 * n / tot, ins, ln, src, big))
if len(sys.argv) > 4:       # line ranges a-b,c-d,...: samples and instructions per range
    for rng in sys.argv[4].split(','):
        a, b = map(int, rng.split('-'))
        ns = sum(v[1] for k, v in per.items() if a <= k <= b); ni = sum(v[2] for k, v in per.items() if a <= k <= b)
        print("lines %d-%d: %.1f%% of the samples, %d warp instructions" % (a, b, 100.0 * ns / tot, ni))
