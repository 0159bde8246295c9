"""ncu -i <rep> --page source --csv --print-source cuda,sass  ->  per source line: share of the warp instructions, lanes per instruction,
share of the stall samples (first captured launch).  usage: ncu_source_lines.py <rep> [top=40] [--sass file:line-from:line-to]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sass = {}; cur = None; hdr = None; f = None; src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":
        cur = (f, int(r[0])); src[cur] = r[1].strip(); continue
    try:
        a = int(r[2], 16)
        if a not in sass: sass[a] = (a, r[3].strip(), int(r[iI]), int(r[iT]), int(r[iS]), cur)
    except (ValueError, IndexError):
        pass
sass = sorted(sass.values())
totI = sum(o[2] for o in sass); totT = sum(o[3] for o in sass); totS = sum(o[4] for o in sass)
print(f"{len(sass)} SASS instructions, {totI:.4g} warp instructions executed, {totT/totI:.2f} lanes per instruction, {totS} samples")
agg = {}
for o in sass:
    a = agg.setdefault(o[5], [0, 0, 0, 0]); a[0] += o[2]; a[1] += o[3]; a[2] += o[4]; a[3] += 1
print(f"{'file':<20}{'line':>5} {'%inst':>6} {'lanes':>6} {'%smpl':>6} {'sass':>5}  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:20]:<20}{k[1]:>5} {100*a[0]/totI:>6.2f} {a[1]/max(a[0],1):>6.1f} {100*a[2]/totS:>6.2f} {a[3]:>5}  {src.get(k, '')[:100]}")
for arg in sys.argv[2:]:
    if arg.startswith("--sass"):
        fn, lo, hi = arg.split("=")[1].split(":")
        sel = [i for i, o in enumerate(sass) if o[5] and o[5][0].startswith(fn) and int(lo) <= o[5][1] <= int(hi)]
        for o in sass[min(sel):max(sel)+1]:
            if o[2]: print(f"{o[0]-sass[0][0]:5x} {o[1][:58]:<58} {100*o[2]/totI:6.3f}% lanes {o[3]/max(o[2],1):4.1f} smpl {o[4]:6d}  {o[5][0][:12]}:{o[5][1]}")
