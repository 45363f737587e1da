"""Joins `ncu --page source --print-source sass` (per-instruction counts) with nvdisasm -g line info to give
instructions executed / stall samples per CUDA source line.  usage: line_profile.py rep.ncu-rep obj.o mangled_fn [kernel_idx]
e.g. TOP=60 python tools/line_profile.py gpurun_out/r01q_full.ncu-rep sarpro_b200/build/kernels_hmma.cu.o \
     _ZN6sarpro6k_hmmaILb1EEEvNS_11HResizeArgsENS_10HMmaParamsE 0
kernel_idx counts the launches in the report (all kernels, in capture order); the object must be the build the report was
captured from (the instruction counts are matched one to one)."""
import csv, re, subprocess, sys, os, tempfile, collections
rep, obj, fn = sys.argv[1:4]; kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur)
    elif cur is not None: cur.append(r)
b = blocks[kidx]; hdr = b[0]; data = b[1:]; ci = {h: i for i, h in enumerate(hdr)}
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + fn + ":"))
lines = []; curline = None
for l in dis[start + 1:]:
    if l.startswith("//-----") : break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(curline)
assert len(lines) == len(data), (len(lines), len(data))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for ln, r in zip(lines, data):
    def f(k):
        try: return float(r[ci[k]])
        except Exception: return 0.0
    a = agg[ln]; a[0] += f("Instructions Executed"); a[1] += f("# Samples"); a[2] += f("L1 Wavefronts Shared")
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
src = {}
for (fnm, n) in agg:
    if fnm and fnm not in src:
        for root in ("sarpro_b200/csrc", "."):
            p = os.path.join(root, fnm)
            if os.path.exists(p): src[fnm] = open(p).read().splitlines(); break
print(f"total warp-inst {tot:.0f} samples {ts:.0f}")
for (k, a) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(os.environ.get("TOP", 40))]:
    text = src.get(k[0], [""] * 100000)[k[1] - 1].strip()[:80] if k and k[0] in src else ""
    print(f"{a[0]/tot*100:5.1f}% inst  {a[1]/ts*100:5.1f}% samp  wf={a[2]/1e6:6.1f}M  {k[0]}:{k[1]}  {text}")
