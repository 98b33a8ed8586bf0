"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py rep [min_share]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; cur = None
agg = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Function Name": fn = r[1][:50]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr, r))
    # rows: CUDA line rows have "Line No"; sass rows have Address.  The 2nd 'Source' column overrides in dict -> use indices
    ln = r[0]
    if ln.strip():
        cur = (fn, int(ln), r[1][:100]); agg.setdefault(cur, [0.0, 0.0, 0.0])
        continue
    if cur is None: continue
    def f(k):
        try: return float(d.get(k, "0").replace(",", "") or 0)
        except ValueError: return 0.0
    a = agg[cur]; a[0] += f("Instructions Executed"); a[1] += f("Warp Stall Sampling (All Samples)"); a[2] += f("Thread Instructions Executed")
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instr {ti:.0f}, samples {ts:.0f}")
for (fn, ln, src), a in agg.items():
    if a[0] / ti >= thr or a[1] / ts >= thr:
        print(f"{ln:4d} inst {a[0]/ti*100:5.1f}%  stall {a[1]/ts*100:5.1f}%  lanes {a[2]/max(a[0],1):4.1f}  {src}")
