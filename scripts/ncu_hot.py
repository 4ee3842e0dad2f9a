"""Top stall-sampled SASS instructions of one kernel from an .ncu-rep (source page, csv).
usage: python scripts/ncu_hot.py <rep> <kernel-regex> [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernel instances may follow each other: split at header rows
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_idx[0]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
h = rows[start]
ci, si = h.index("# Samples"), h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[start + 1:end]:
    if len(r) <= ci:
        continue
    try:
        v = float(r[ci])
    except ValueError:
        continue
    top = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    data.append((v, r[si][:90], ",".join(f"{n[6:]}:{int(x)}" for x, n in top if x > 0)))
tot = sum(v for v, _, _ in data) or 1
print(f"{kern}: {len(data)} instrs, {int(tot)} samples")
for i, (v, s, t) in enumerate(data):
    data[i] = (v, i, s, t)
for v, i, s, t in sorted(data, reverse=True)[:N]:
    print(f"{v / tot * 100:5.1f}%  #{i:<5d} {s:90s} {t}")
