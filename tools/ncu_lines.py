"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
The source page lists every instruction under the line of each file it is attributed to (the inlined callee's line AND the
call site's line), so shares are taken within ONE file section: pass the file that holds the kernel body.
    python tools/ncu_lines.py rep [min_share] [function substring] [file substring]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
fn_f = sys.argv[3] if len(sys.argv) > 3 else ""
file_f = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; cur = None; fn = ""; path = ""
agg = collections.OrderedDict()
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path": path = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if fn_f not in fn or file_f not in path: continue
    if r[0].strip():
        cur = (fn[:40], path.split("/")[-1], int(r[0]), r[1][:100]); agg.setdefault(cur, [0.0, 0.0, 0.0])
        continue
    if cur is None: continue
    d = dict(zip(hdr, r))
    def f(k):
        try: return float(d.get(k, "0").replace(",", "") or 0)
        except ValueError: return 0.0
    a = agg[cur]; a[0] += f("Instructions Executed"); a[1] += f("Warp Stall Sampling (All Samples)"); a[2] += f("Thread Instructions Executed")
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instr {ti:.0f}, samples {ts:.0f}")
for (fn, fl, ln, src), a in agg.items():
    if a[0] / ti >= thr or a[1] / ts >= thr:
        print(f"{fl[:16]:16s} {ln:4d} inst {a[0]/ti*100:5.1f}%  stall {a[1]/ts*100:5.1f}%  lanes {a[2]/max(a[0],1):4.1f}  {src}")
