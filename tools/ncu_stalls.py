"""Top stall locations of one kernel from an ncu report (needs -lineinfo + --import-source on).
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel regex> [N]
"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print(rows[0][1][:90], "total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {k: sum(f(r, k) for r in data) for k in stalls}
print("by reason:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:N]:
    s = sorted(((f(r, k), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{f(r,'# Samples'):7.0f} {100*f(r,'# Samples')/max(tot,1):5.1f}%  {r[idx['Source']].strip()[:64]:64s} {s[0][1]}={s[0][0]:.0f} {s[1][1]}={s[1][0]:.0f}")
