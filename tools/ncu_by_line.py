"""Join an ncu SASS source page (csv) with nvdisasm line info and aggregate executed
instructions per source line / per estimator block.
usage: python tools/ncu_by_line.py <prof.ncu-rep | exported source page .csv> <lib.so> <mangled-kernel-substring>
(the .csv form = `ncu -i prof.ncu-rep --page source --csv`, what tools/ncu_export.sh brings back from the GPU box)"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "synth" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
addr2line = {}
infn = False
cur = ("?", 0)
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
if rep.endswith(".csv"):
    src = open(rep).read().splitlines()
else:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(src))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isamp = hdr.index("# Samples")
base = None
per_line = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) <= it or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    key = addr2line.get(a - base, (("?", 0), ""))[0]
    e, t, s = int(float(r[ie] or 0)), int(float(r[it] or 0)), int(float(r[isamp] or 0))
    per_line[key][0] += e; per_line[key][1] += t; per_line[key][2] += s
    tot[0] += e; tot[1] += t; tot[2] += s
print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} avg threads {tot[1]/max(tot[0],1):.2f} samples {tot[2]}")
print("top source lines by warp instructions:")
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"  {k[0]}:{k[1]:<5d} warp-inst {v[0]/tot[0]*100:5.1f}%  avg-thr {v[1]/max(v[0],1):5.1f}  stall-samples {v[2]/max(tot[2],1)*100:5.1f}%")

# ---- coarse categories (source line ranges of this revision are passed as env CVR_CATS or default)
import json
cats_file = os.environ.get("CVR_CATS")
if cats_file:
    cats = json.load(open(cats_file))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (f, l), v in per_line.items():
        name = "other"
        for c in cats:
            if c["file"] == f and c["lo"] <= l <= c["hi"]:
                name = c["name"]; break
        for i in range(3): agg[name][i] += v[i]
    print("by block:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:<22s} warp-inst {v[0]/tot[0]*100:5.1f}%  avg-thr {v[1]/max(v[0],1):5.1f}  thread-inst {v[1]/tot[1]*100:5.1f}%  stall-samples {v[2]/max(tot[2],1)*100:5.1f}%")
