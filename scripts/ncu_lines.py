"""Attributes executed warp instructions to CUDA source lines (needs -lineinfo): ncu_lines.py rep warps steps"""
import csv, collections, subprocess, sys
rep, warps, steps = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}; ops = collections.Counter(); tot = 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] == 'Line No': continue
    if r[0] != '':
        key = (cur, int(r[0]), r[1].strip()[:90]); per[key] = per.get(key, 0) + (int(r[7]) if r[7].isdigit() else 0)
    elif r[3] not in ('-', '') and r[7] not in ('-', ''):
        t = r[3].split(); op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0] if not op.startswith('IMAD') else op] += int(r[7]); tot += int(r[7])
print("warp-instr per warp-step: %.1f" % (tot / (warps * steps)))
print(' '.join('%s=%.1f' % (k, v / (warps * steps)) for k, v in ops.most_common(22)))
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print("%6.1f  %s:%d  %s" % (v / (warps * steps), k[0], k[1], k[2]))
