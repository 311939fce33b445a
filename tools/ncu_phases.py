"""Per-phase breakdown of one kernel from `ncu -i rep --page source --csv`: the SASS is split at
every BAR.SYNC; per segment: warp-stall samples, instructions, shared-memory wavefronts (total /
ideal), and the top shared-memory instructions by excess wavefronts.
usage: python tools/ncu_phases.py report.ncu-rep"""
import csv, subprocess, sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
segs, cur = [], []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    cur.append(r)
    if "BAR.SYNC" in r[col["Source"]]:
        segs.append(cur); cur = []
if cur: segs.append(cur)
tot_s = sum(num(r, "# Samples") for s in segs for r in s)
tot_w = sum(num(r, "L1 Wavefronts Shared") for s in segs for r in s)
print(f"total samples {tot_s:.0f}, shared wavefronts {tot_w/1e6:.1f} M")
for i, s in enumerate(segs):
    smp = sum(num(r, "# Samples") for r in s)
    ins = sum(num(r, "Instructions Executed") for r in s)
    w = sum(num(r, "L1 Wavefronts Shared") for r in s)
    wi = sum(num(r, "L1 Wavefronts Shared Ideal") for r in s)
    stalls = {k: sum(num(r, k) for r in s) for k in ("stall_barrier", "stall_short_sb", "stall_long_sb", "stall_mio", "stall_wait", "stall_math", "stall_not_selected", "stall_selected")}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
    print(f"segment {i}: {len(s)} SASS lines, samples {100*smp/tot_s:.1f}%, inst {ins/1e6:.1f} M, wavefronts {w/1e6:.1f} M (ideal {wi/1e6:.1f} M)  "
          + " ".join(f"{k[6:]}={100*v/max(smp,1):.0f}%" for k, v in top))
    worst = sorted(s, key=lambda r: -(num(r, "L1 Wavefronts Shared") - num(r, "L1 Wavefronts Shared Ideal")))[:6]
    for r in worst:
        ex = num(r, "L1 Wavefronts Shared") - num(r, "L1 Wavefronts Shared Ideal")
        if ex < 2e5: continue
        print(f"      +{ex/1e6:.2f} M of {num(r,'L1 Wavefronts Shared')/1e6:.2f} M  x{num(r,'Instructions Executed')/1e6:.2f} M  {r[col['Source']].strip()[:70]}")
