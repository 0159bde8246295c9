"""regression/0010_pt at the reference's arguments through the corona_b200 command line: progressions per call chosen by the
binary (no --batch), one (--batch 1, the reference's default grouping) and 16.  Numbers go to profiles/README.md."""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage
REF = os.path.join(ROOT, "oracle", "_ref")
g = GoldenImage("c10")
nra2 = g.write_files(tempfile.mkdtemp())
for sampler in ("pt", "ptdl"):
    for extra in ([], ["--batch", "1"], ["--batch", "16"]):
        p = subprocess.run([os.path.join(ROOT, "corona-13_b200", "corona_b200"), nra2, "-s", "512", "-w", "1024", "-h", "576", "--frame", "1", *extra,
                            "--sampler", sampler, "--points", "rand", "--coeff", os.path.join(REF, "data", "ergb2spec.coeff"),
                            "--tables", os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")], capture_output=True, text=True)
        f = float(re.findall(r"average of ([0-9.]+) s/frame", p.stdout)[0])
        print(sampler, " ".join(extra) or "auto", f"{f*1e3:.3f} ms/progression", f"{1/f:.0f} spp/s", flush=True)
