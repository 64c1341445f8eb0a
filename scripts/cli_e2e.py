"""End-to-end drop-in run of the gemma-b200 CLI on a synthetic PLINK data set (GPU box): -gk, then -lmm 4 with the K it wrote.
Prints one JSON line with the wall time of every phase as reported in the CLI's own log plus the total process times.
usage: python scripts/cli_e2e.py [n] [p] [outdir]"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gemma_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/cli_e2e"
os.makedirs(out, exist_ok=True)
prefix = os.path.join(out, "syn")
cli = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")

t0 = time.time()
bed = synth.make_bed_torch(n, p, torch.device("cuda", 0), seed=2024, miss_rate=0.002).cpu().numpy()
with open(prefix + ".bed", "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01])); f.write(bed.tobytes())
g64 = synth.genotypes(n, 64, seed=2024, snp_offset=0, miss_rate=0.0).astype(np.float64)
y = synth.phenotype(n, g64, 7)
with open(prefix + ".fam", "w") as f:
    for i in range(n):
        f.write("f%d i%d 0 0 1 %.6f\n" % (i, i, y[i]))
with open(prefix + ".bim", "w") as f:
    for s in range(p):
        f.write("%d\trs%d\t0\t%d\tA\tG\n" % (1 + s * 22 // p, s, 1000 + s))
t_gen = time.time() - t0


def run(args):
    t = time.time()
    r = subprocess.run([cli] + args, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(1)
    return time.time() - t, r.stdout


t_gk, o1 = run(["-bfile", prefix, "-gk", "1", "-bin", "-o", "k", "-outdir", out])
t_lmm_txt, o2 = run(["-bfile", prefix, "-k", out + "/k.cXX.txt", "-lmm", "4", "-o", "a_txt", "-outdir", out])
t_lmm_bin, o3 = run(["-bfile", prefix, "-k", out + "/k.cXX.txt.bin", "-lmm", "4", "-o", "a_bin", "-outdir", out])


def logtimes(name):
    d = {}
    for ln in open(os.path.join(out, name + ".log.txt")):
        if "time" in ln and "=" in ln:
            k, v = ln.strip("# \n").split("=")
            try:
                d[k.strip()] = float(v.split()[0])
            except ValueError:
                pass
    return d


a = [l.split("\t") for l in open(out + "/a_txt.assoc.txt").read().splitlines()[1:]]
b = [l.split("\t") for l in open(out + "/a_bin.assoc.txt").read().splitlines()[1:]]
pa = np.array([[float(x) for x in r[7:]] for r in a]); pb = np.array([[float(x) for x in r[7:]] for r in b])
ok = np.isfinite(pa) & np.isfinite(pb)
print(json.dumps({"n": n, "p": p, "analysed_snps": len(a), "generate_s": round(t_gen, 1),
                  "gk_wall_s": round(t_gk, 2), "lmm4_wall_s_text_K": round(t_lmm_txt, 2), "lmm4_wall_s_bin_K": round(t_lmm_bin, 2),
                  "gk_log_min": logtimes("k"), "lmm_log_min": logtimes("a_bin"),
                  "snps_per_s_whole_process": round(len(a) / t_lmm_bin, 1),
                  "text_vs_bin_K_max_rel_diff": float(np.max(np.abs(pa[ok] - pb[ok]) / np.maximum(np.abs(pb[ok]), 1e-300))),
                  "K_text_bytes": os.path.getsize(out + "/k.cXX.txt"), "K_bin_bytes": os.path.getsize(out + "/k.cXX.txt.bin")}))
