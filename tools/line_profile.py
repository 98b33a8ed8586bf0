"""Attribute an ncu --set full capture of k_layer to source lines / kernel phases.
   usage: python tools/line_profile.py gpurun_out/prof.ncu-rep [mangled-kernel-substring]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "_Z7k_layerILi0ELi0ELb0ELb1EE"
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {ROOT}/voxel-stack-blender_b200/libvsb_b200.so > /dev/null 2>&1 && nvdisasm -g vsb_layer.sm_100a.cubin > all.txt 2>/dev/null", shell=True)
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
blocks, cur = [], None
want = "k_layer" if "k_layer" in kname else kname
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append((r[1] if len(r) > 1 else "", cur)); continue
    if cur is not None:
        cur.append(r)
b = [blk for name, blk in blocks if want in name][-1]; hdr = b[0]; data = b[1:]   # last matching launch in the report
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
txt = open(os.path.join(tmp, "all.txt")).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + kname) and l.endswith(":")][0]
lines, curline = [], None
for l in txt[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.search(r'//## File ".*vsb_(\w+)\.(cuh?)", line (\d+)', l)
    if m:
        curline = (m.group(1), int(m.group(3))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(curline)
assert len(lines) == len(data), (len(lines), len(data))
layer = open(f"{ROOT}/voxel-stack-blender_b200/csrc/vsb_layer.cu").read().split("\n")
marks = [(i + 1, l.strip()) for i, l in enumerate(layer) if re.search(r"// ---- (P\d|once per CTA)", l) or "if (hb > 0) {" in l and "total" in layer[i + 1]]
def phase(f, l):
    if f == "tile": return "search (tile.cuh)"
    if f == "common": return "LUT map / helpers"
    if f != "layer": return str(f)
    name = "entry/setup"
    for ln, text in marks:
        if l >= ln: name = re.sub(r"[-/ ]+$", "", text.replace("// ----", "").strip())[:40] or name
    return name
agg, smp, lagg, lsmp = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for i, ln in enumerate(lines):
    f, l = ln if ln else ("?", 0)
    k = phase(f, l)
    agg[k] += int(data[i][ie]); smp[k] += int(data[i][ss]); lagg[(f, l)] += int(data[i][ie]); lsmp[(f, l)] += int(data[i][ss])
tot, tots = sum(agg.values()), sum(smp.values())
print(f"total warp-instructions {tot}, samples {tots}")
for k, c in smp.most_common():
    print(f"{k:44s} instr {100*agg[k]/tot:5.1f}%   samples {100*c/tots:5.1f}%")
print("--- top lines by samples")
srcs = {"layer": layer, "tile": open(f"{ROOT}/voxel-stack-blender_b200/csrc/vsb_tile.cuh").read().split("\n"),
        "common": open(f"{ROOT}/voxel-stack-blender_b200/csrc/vsb_common.cuh").read().split("\n")}
for (f, l), c in (sorted(lagg.items(), key=lambda kv: -kv[1])[:int(os.environ.get("TOPN","28"))] if os.environ.get("BYINSTR") else lsmp.most_common(int(os.environ.get("TOPN","28")))):
    c = lsmp[(f, l)]
    t = srcs.get(f, [""])[l - 1].strip()[:88] if f in srcs and 0 < l <= len(srcs[f]) else ""
    print(f"{100*c/tots:5.1f}% smp {100*lagg[(f,l)]/tot:5.1f}% instr  {f}:{l:4}  {t}")
