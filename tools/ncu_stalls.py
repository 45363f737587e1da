"""Per-instruction stall summary of an .ncu-rep (source page): stall reasons overall, hottest instructions, opcode mix."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
b = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0] not in ("Kernel Name", "Address")]
tot = collections.Counter(); samples = 0
for r in b:
    for s in stall_cols:
        try: tot[s] += int(r[idx[s]])
        except Exception: pass
    samples += int(r[idx["# Samples"]])
print("samples", samples)
for s, v in tot.most_common(12): print(f"  {s:28s} {v:8d} {100*v/samples:5.1f}%")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for r in sorted(b, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    st = sorted(((s, int(r[idx[s]])) for s in stall_cols if int(r[idx[s]]) > 0), key=lambda x: -x[1])[:3]
    print(r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[1].strip()[:72].ljust(72), st)
op = collections.Counter(); ops = collections.Counter()
for r in b:
    o = [x for x in r[1].split() if not x.startswith("@")][0].split(".")[0]
    op[o] += int(r[idx["Instructions Executed"]]); ops[o] += int(r[idx["# Samples"]])
tt = sum(op.values())
print("executed warp instructions", tt)
for o, v in op.most_common(22): print(f"  {o:12s} {v:10d} {100*v/tt:5.1f}%  samples {100*ops[o]/samples:5.1f}%")
