"""Hot source lines of an ncu --set full --import-source on report: warp-stall samples aggregated per CUDA source line.
    python tools/ncu_hot.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.Counter(); text = {}; stalls = collections.defaultdict(collections.Counter)
fname = "?"; hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; k = hdr.index("Warp Stall Sampling (All Samples)"); st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) <= k:
        continue
    try:
        v = float(r[k])
    except ValueError:
        continue
    key = (fname, r[0])
    agg[key] += v; text[key] = r[1]
    for i in st:
        try:
            stalls[key][hdr[i]] += float(r[i])
        except ValueError:
            pass
tot = sum(agg.values()) or 1
print("total samples", tot)
for key, v in agg.most_common(top):
    s = ", ".join(f"{n[6:]} {int(c)}" for n, c in stalls[key].most_common(3) if c > 0)
    print(f"{100 * v / tot:5.1f}%  {key[0]}:{key[1]:>4}  {text[key].strip()[:110]}   [{s}]")
