"""Reads `ncu --page source --print-source sass --csv` of one kernel and prints the hot regions:
instructions executed, stall samples by reason, shared wavefronts, per SASS line (top-N) and totals."""
import csv, sys, subprocess
rep = sys.argv[1]; kern_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# split kernels
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(r)
b = blocks[kern_idx]
hdr = b["rows"][0]; data = b["rows"][1:]
ci = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
tot_wf = sum(f(r, "L1 Wavefronts Shared") for r in data)
print(b["name"][:100]); print(f"SASS lines {len(data)} warp-inst {tot_inst:.0f} samples {tot_samp:.0f} smem wavefronts {tot_wf:.0f} (ideal {sum(f(r,'L1 Wavefronts Shared Ideal') for r in data):.0f})")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v > 0))
# opcode histogram weighted by executions
ops = {}
for r in data:
    src = r[ci["Source"]].strip(); op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + f(r, "Instructions Executed")
print("executed by opcode:", ", ".join(f"{k}={v/tot_inst*100:.1f}%" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:25]))
print("--- top lines by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{r[ci['Address']][-5:]} samp={f(r,'# Samples'):6.0f} inst={f(r,'Instructions Executed'):9.0f} wf={f(r,'L1 Wavefronts Shared'):9.0f} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f} | {r[ci['Source']][:70]}")
