"""Per-source-line instruction counts and stall samples of an .ncu-rep (compiled with -lineinfo, captured with
--import-source on).  usage: python tools/ncu_source_hot.py report.ncu-rep [kernel-substring] [top-N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = None; hdr = None; agg = {}
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0] == "File Path": fpath = row[1]; continue
    if row[0] == "Function Name": fn = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or fn is None or filt not in fn: continue
    if row[0] == "": continue          # SASS rows
    try:
        ie = hdr.index("Instructions Executed"); ws = hdr.index("Warp Stall Sampling (All Samples)")
        key = (fn.split("(")[0], fpath.split("/")[-1], int(row[0]))
        a = agg.setdefault(key, [0, 0, row[1]])
        a[0] += int(row[ie]); a[1] += int(row[ws])
    except (ValueError, IndexError):
        pass
by_fn = {}
for (f, path, ln), (ie, ws, src) in agg.items():
    by_fn.setdefault(f, []).append((ie, ws, path, ln, src))
for f, rows in by_fn.items():
    tot = sum(r[0] for r in rows); tots = sum(r[1] for r in rows)
    print(f"== {f}: {tot:,} warp instructions, {tots:,} stall samples")
    for ie, ws, path, ln, src in sorted(rows, reverse=True)[:top]:
        print(f"  {100*ie/max(tot,1):5.1f}% inst {100*ws/max(tots,1):5.1f}% stall  {path}:{ln}: {src.strip()[:110]}")
