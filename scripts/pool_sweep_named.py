"""scratch: regression/0010_pt geometry at 1024x576 through the corona_b200 command line, path pool size (CB200_POOL_PATHS) x
progressions per call (--batch).  Numbers go to profiles/."""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage
REF = os.path.join(ROOT, "oracle", "_ref")
g = GoldenImage("c10")
nra2 = g.write_files(tempfile.mkdtemp())
samplers = sys.argv[1].split(",") if len(sys.argv) > 1 else ["ptdl"]
for sampler in samplers:
    for pool in ("2097152", "", "8388608", "16777216"):
        for batch in ("4", "16", "32", "64"):
            env = dict(os.environ)
            if pool: env["CB200_POOL_PATHS"] = pool
            best = 1e9
            for rep in range(2):
                p = subprocess.run([os.path.join(ROOT, "corona-13_b200", "corona_b200"), nra2, "-s", "1024", "-w", "1024", "-h", "576", "--frame", "1", "--batch", batch,
                                    "--sampler", sampler, "--points", "rand", "--coeff", os.path.join(REF, "data", "ergb2spec.coeff"),
                                    "--tables", os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")], capture_output=True, text=True, env=env)
                m = re.findall(r"average of ([0-9.]+) s/frame", p.stdout)
                if not m: print("FAILED", p.stderr[-300:]); break
                best = min(best, float(m[0]))
            print(sampler, "pool", pool or "default", "batch", batch, f"{best*1e3:.3f} ms/progression", f"{1/best:.0f} spp/s", flush=True)
