"""Summary table of an `ncu --page raw --csv` export (made on the GPU box by scripts/gpu/r02_profiles.sh): one line per
distinct (kernel, grid) with the last captured launch's time, DRAM bytes, registers, occupancy and pipe utilisation.
usage: ncu_raw_table.py file.raw.csv [peak_GBps]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6554.2
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try:
        return float(r[ix[k]])
    except Exception:
        return d
seen = {}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    seen[(r[ix['Kernel Name']], r[ix['launch__grid_size']])] = r
units = rows[1]
tu = units[ix['gpu__time_duration.sum']]; bu = units[ix['dram__bytes_read.sum']]
tscale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(tu, 1.0)
bscale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(bu, 1.0)
print("%-58s %7s %5s %5s %9s %9s %9s %6s %6s %6s %6s" % ("kernel", "grid", "blk", "regs", "time_us", "dram_rd_MB", "dram_wr_MB", "GB/s", "%peak", "issue%", "fp64%"))
for (name, grid), r in seen.items():
    t = g(r, 'gpu__time_duration.sum') * tscale
    rd = g(r, 'dram__bytes_read.sum') * bscale; wr = g(r, 'dram__bytes_write.sum') * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[ix['dram__bytes_write.sum']], 1.0)
    gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0.0
    print("%-58s %7s %5s %5s %9.1f %9.2f %9.2f %6.0f %6.1f %6.1f %6.1f" % (
        name[:58], grid, r[ix['launch__block_size']], r[ix['launch__registers_per_thread']], t, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak,
        g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')))
