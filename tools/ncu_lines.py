"""Per-source-line stall-sample summary of an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows[:20]) if '# Samples' in r)
hdr = rows[hi]; ni = hdr.index('# Samples')
keys = ['stall_barrier', 'stall_long_sb', 'stall_math', 'stall_wait', 'stall_short_sb', 'stall_mio', 'stall_lg',
        'stall_branch_resolving', 'stall_not_selected', 'stall_selected', 'stall_no_inst', 'stall_dispatch']
cols = {k: hdr.index(k) for k in keys}
lines = []
for r in rows[hi + 1:]:
    if r and r[0].strip().isdigit() and len(r) > ni and r[ni] not in ('', '-'):
        d = {k: (int(r[c]) if r[c] not in ('-', '') else 0) for k, c in cols.items()}
        lines.append((int(r[0]), r[1].strip(), int(r[ni]), d))
tot = sum(l[2] for l in lines)
print("total samples", tot)
agg = {k: sum(l[3][k] for l in lines) for k in keys}
print("by reason:", ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))
for ln, src, n, d in sorted(lines, key=lambda l: -l[2])[:top]:
    t3 = ", ".join("%s:%d%%" % (k.replace('stall_', ''), 100 * v // max(n, 1)) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3])
    print("%4d %5.1f%%  %-78s | %s" % (ln, 100.0 * n / tot, src[:78], t3))
