"""Per-role breakdown of the warp-specialised RING kernel from `ncu -i rep --page source --csv`: the SASS is split where
the job warps' code ends (the first EXIT after the last MUFU.RCP64H of the job loops); per role: warp samples (taken at
speed), instructions (hardware count, smsp__inst_executed; the per-line counts of the source page come from an
instrumented replay and overstate the barrier wait loops), shared-memory wavefronts, stall reasons, hot lines; then the
wait loops of ring_mbar_wait.      usage: python tools/ncu_roles.py report.ncu-rep"""
import csv, subprocess, sys

rep = sys.argv[1]
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout.splitlines()))
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout.splitlines()))
R = dict(zip(raw[0], raw[2]))
hdr, data = src[1], [r for r in src[2:] if len(r) >= len(src[1])]
col = {}
for i, h in enumerate(hdr):
    col.setdefault(h, i)
def num(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
S = lambda r: r[col["Source"]]
first_out = len(data)
rcp = [i for i, r in enumerate(data) if "MUFU.RCP64H" in S(r)]
jobs_rcp = [i for i in rcp if i < len(data) // 2]
if jobs_rcp:
    for i in range(jobs_rcp[-1], len(data)):
        if "EXIT" in S(data[i]):
            first_out = i + 1
            break
tot = sum(num(r, "# Samples") for r in data)
print(f"{src[0][1] if len(src[0]) > 1 else 'kernel'}: {len(data)} SASS lines, {tot:.0f} warp samples, "
      f"{float(R['gpu__time_duration.sum']):.1f} us under ncu, {float(R['smsp__inst_executed.sum'])/1e6:.1f} M warp instructions")
print(f"LSU data pipe {float(R['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']):.1f} % of peak "
      f"(shared memory {float(R['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'])/1e6:.1f} M wavefronts = "
      f"{float(R['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']):.1f} %), issue slots "
      f"{float(R['smsp__issue_active.avg.pct_of_peak_sustained_active']):.1f} %, FP64 pipe {float(R['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']):.1f} %, "
      f"DRAM {float(R['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):.1f} %")
stalls = [h for h in col if h.startswith("stall_") and "Not Issued" not in h]
allst = {k: sum(num(r, k) for r in data) for k in stalls}
print("stall reasons, share of all warp samples: " + ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:9]))
for name, seg in (("prologue + JOB warps", data[:first_out]), ("WRITE-OUT warps", data[first_out:])):
    smp = sum(num(r, "# Samples") for r in seg)
    w = sum(num(r, "L1 Wavefronts Shared") for r in seg)
    st = {k: sum(num(r, k) for r in seg) for k in stalls}
    print(f"{name}: {len(seg)} SASS lines, samples {100*smp/tot:.1f}%, issued (selected samples) {100*st.get('stall_selected',0)/max(allst.get('stall_selected',1),1):.1f}% of all issues, "
          f"shared wavefronts by instruction {w/1e6:.1f} M | " + " ".join(f"{k[6:]}={100*v/max(smp,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5]))
base = 0
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:8]:
    i = data.index(r)
    top = sorted(((k, num(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f"  hot {i:5d} {100*num(r,'# Samples')/tot:5.2f}%  {S(r).strip()[:58]:58s} " + " ".join(f"{k[6:]}={v:.0f}" for k, v in top))
inloop, last = [], None
for i, r in enumerate(data):
    s = S(r)
    if "SYNCS.PHASECHK" in s or "NANOSLEEP" in s or ("BRA" in s and last is not None and i - last <= 2):
        inloop.append(i); last = i
sel = sum(num(data[i], "stall_selected") for i in inloop)
print(f"barrier wait loops (ring_mbar_wait): {100*sum(num(data[i],'# Samples') for i in inloop)/tot:.1f}% of the warp samples "
      f"({100*sum(num(data[i],'stall_long_sb') for i in inloop)/tot:.1f}% asleep), {100*sel/max(allst.get('stall_selected',1),1):.1f}% of the issue slots")
