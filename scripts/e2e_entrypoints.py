"""End-to-end run of the entry points (src/simulate/simulate.py, src/genofix.py) on a small synthetic case."""
import os, subprocess, sys, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genofix_b200 import synth
tmp = os.path.join(ROOT, "gpurun_out", "e2e_tmp")
os.makedirs(tmp, exist_ok=True)
rows = synth.make_pedigree(800, 6, sire_frac=0.1, dam_frac=0.5, seed=3)
np.savetxt(os.path.join(tmp, "ped.fam"), rows, fmt="%d")
with open(os.path.join(tmp, "snps.map"), "w") as f:
    for i in range(1500):
        f.write("1 SNP%d %.2f %d\n" % (i + 1, 0.01 * i, 1000 * i))
r = subprocess.run([sys.executable, "src/simulate/simulate.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/simout",
                    "-c", "500", "-t", "0.4", "-l", "0.64"], capture_output=True, text=True, cwd=ROOT)
print(r.stdout[-900:]); print(r.stderr[-600:])
r = subprocess.run([sys.executable, "src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/gfout",
                    "-i", tmp + "/simout/simulatedgenome_with_1errors.ssv.gz", "-C", "600"], capture_output=True, text=True, cwd=ROOT)
print(r.stdout[-700:]); print(r.stderr[-600:])
r = subprocess.run([sys.executable, "src/simulate/simulate.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/simdev",
                    "-c", "500", "-t", "0.4", "-l", "0.64", "--device_sim"], capture_output=True, text=True, cwd=ROOT)
print("device_sim:", r.stdout[-600:]); print(r.stderr[-600:])
# the same matrix through the 2-bit container
from genofix_b200.gfutils.packedgens import PackedGens
PackedGens.from_text(tmp + "/simout/simulatedgenome_with_1errors.ssv.gz", tmp + "/errors.g2b")
r = subprocess.run([sys.executable, "src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/gfout2",
                    "-i", tmp + "/errors.g2b", "-C", "600"], capture_output=True, text=True, cwd=ROOT)
print("g2b:", r.stdout[-300:]); print(r.stderr[-600:])
import filecmp
print("g2b output identical to text-input output:", filecmp.cmp(tmp + "/gfout/simulatedgenome_corrected_errors_threshold.ssv",
                                                                tmp + "/gfout2/simulatedgenome_corrected_errors_threshold.ssv", shallow=False))
r = subprocess.run([sys.executable, "src/gfutils/extract_founders.py", "-p", tmp + "/ped.fam"], capture_output=True, text=True, cwd=ROOT)
print("founders:", len(r.stdout.split()), r.stderr[-300:])
shutil.rmtree(tmp, ignore_errors=True)
