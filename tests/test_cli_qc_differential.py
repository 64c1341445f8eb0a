"""Differential test of the host side of the CLI (readers + individual / SNP QC) against the reference's
OWN CLI (oracle/_ref/gemma_ref, the unmodified src/*.cpp) on randomised BIMBAM and PLINK inputs with the awkward cases mixed in:
NA phenotypes / covariates, monomorphic SNPs, dosage-valued genotypes, mixed separators, a shuffled and incomplete annotation file, SNPs collinear with a covariate, and the
-miss / -maf / -hwe / -r2 / -notsnp / -n / -nind / -gxe / -widv / -loco switches (refusals included).  The reference runs `-lm 1` (cheap, no kinship needed) and its assoc file lists
the analysed SNPs with n_miss and af; `gemma-b200 -qc-only` must select the same SNPs / individuals and print the same counts.
CPU only."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import ref as REF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")
COUNT_KEYS = ("number of total individuals", "number of analyzed individuals", "number of covariates", "number of total SNPs/var",
              "number of analyzed SNPs")


def _make_case(d, seed, plink, n=60, l=120):
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.0, 0.5, l)
    f[rng.random(l) < 0.1] = 0.01
    G = rng.binomial(2, f[:, None], size=(l, n)).astype(float)
    G[0], G[1], G[2] = 1.0, 0.0, 2.0                                     # monomorphic rows
    miss = rng.random((l, n)) < rng.choice([0.0, 0.02, 0.05, 0.08], size=l)[:, None]
    if not plink:
        dos = rng.random((l, n)) < 0.3                                   # dosage-valued entries (BIMBAM only)
        G = np.where(dos, np.round(np.clip(G + rng.normal(0, 0.2, G.shape), 0, 2), 3), G)
    ph, ph2 = rng.normal(size=n), rng.normal(size=n)
    phna, ph2na = rng.random(n) < 0.1, rng.random(n) < 0.15
    cv = rng.normal(size=(n, 2))
    G[5] = np.clip(np.round(1 + cv[:, 0]), 0, 2)                         # collinear with a covariate: the -r2 filter
    G[6] = np.clip(np.round(1 + 0.8 * cv[:, 1] + 0.3 * rng.normal(size=n)), 0, 2)
    cvna = rng.random(n) < 0.05
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "pheno.txt"), "w") as fo:
        for i in range(n):
            fo.write(("NA" if phna[i] else "%.5f" % ph[i]) + "\t" + ("NA" if ph2na[i] else "%.5e" % ph2[i]) + "\n")
    with open(os.path.join(d, "cvt.txt"), "w") as fo:
        for i in range(n):
            fo.write("1\t" + ("NA" if cvna[i] else "%.4f" % cv[i, 0]) + "\t%.4f\n" % cv[i, 1])
    with open(os.path.join(d, "gxe.txt"), "w") as fo:                      # -gxe with -lm: only drops the individuals without a value
        for i in range(n):
            fo.write(("NA" if rng.random() < 0.08 else "%.3f" % rng.normal()) + "\n")
    with open(os.path.join(d, "w.txt"), "w") as fo:                        # -widv: individuals without a weight are dropped
        for i in range(n):
            fo.write(("NA" if rng.random() < 0.07 else "%.3f" % rng.uniform(0.2, 3.0)) + "\n")
    with open(os.path.join(d, "anno.txt"), "w") as fo:
        for s in rng.permutation(l):                                       # shuffled, incomplete, mixed separators, non-numeric chr
            if s % 11 == 7:
                continue
            sep = [", ", "\t", " ", ","][s % 4]
            fo.write(sep.join(["rs%d" % s, "%d" % (1000 + 10 * s), ["1", "2", "X", "19"][s % 4]]) + ("" if s % 5 else sep + "0.5") + "\n")
    if not plink:
        with gzip.open(os.path.join(d, "geno.txt.gz"), "wt") as fo:
            for s in range(l):
                vals = ["NA" if miss[s, i] else ("%g" % G[s, i]) for i in range(n)]
                sep = ", " if s % 3 else ("\t" if s % 2 else " ")
                fo.write(sep.join(["rs%d" % s, "A", "G"] + vals) + "\n")
        return ["-g", os.path.join(d, "geno.txt.gz"), "-p", os.path.join(d, "pheno.txt"), "-a", os.path.join(d, "anno.txt")]
    Gi = np.where(miss, -1, G).astype(np.int64)
    bed = np.zeros((l, (n + 3) // 4), dtype=np.uint8)
    code = {0: 3, 1: 2, 2: 0, -1: 1}                                     # PLINK: 00 hom A1, 01 missing, 10 het, 11 hom A2
    for s in range(l):
        for i in range(n):
            bed[s, i >> 2] |= code[int(Gi[s, i])] << (2 * (i & 3))
    with open(os.path.join(d, "pl.bed"), "wb") as fo:
        fo.write(bytes([0x6C, 0x1B, 0x01])); fo.write(bed.tobytes())
    with open(os.path.join(d, "pl.bim"), "w") as fo:
        for s in range(l):
            fo.write("%d\trs%d\t0\t%d\tA\tG\n" % (1 + s % 3, s, 1000 + 10 * s))
    with open(os.path.join(d, "pl.fam"), "w") as fo:
        for i in range(n):
            fo.write("f%d i%d 0 0 1 %s\n" % (i, i, "-9" if phna[i] else "%.5f" % ph[i]))
    return ["-bfile", os.path.join(d, "pl")]


def _first_error(txt):
    m = re.findall(r"(?:error!|ERROR:)[^\n]*", txt)
    assert m, txt[-400:]
    return m[0].replace("ERROR: Enforce failed for ", "").split(" in /root")[0].replace("error! ", "").strip()


def _count(txt, key):
    for ln in txt.splitlines():
        if key in ln:
            return ln.split("=")[-1].strip()
    return None


VARIANTS = [[], ["cvt"], ["-maf", "0.05"], ["-miss", "0.03"], ["-hwe", "0.5"], ["cvt", "-r2", "0.3"], ["-maf", "0", "-miss", "0.2"],
            ["-notsnp"], ["cvt", "-hwe", "0.9", "-maf", "0.1"], ["-n", "2"],
            ["-gxe", "GXE"], ["cvt", "-gxe", "GXE", "-maf", "0.05"], ["-gxe", "GXE", "-nind", "30"], ["-loco", "1"],
            ["-widv", "WIDV"], ["cvt", "-widv", "WIDV", "-gxe", "GXE"], ["-widv", "WIDV", "-nind", "30"],
            ["-nind", "7"], ["-nind", "40"], ["-nind", "1000"], ["cvt", "-nind", "45"], ["cvt", "-nind", "1000"]]


@pytest.mark.parametrize("plink", [False, True], ids=["bimbam", "plink"])
@pytest.mark.parametrize("seed", [0, 1])
def test_qc_selection_matches_the_reference_cli(tmp_path, seed, plink):
    if not os.path.exists(REF.EXE) and not os.path.isdir(REF.REF_SRC):
        pytest.skip("reference CLI not built")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gemma_b200", "host")])
    d = str(tmp_path)
    base = _make_case(d, seed, plink)
    n_refused = 0
    for k, v in enumerate(VARIANTS):
        args = base + (["-c", os.path.join(d, "cvt.txt")] if "cvt" in v else []) + [{"GXE": os.path.join(d, "gxe.txt"), "WIDV": os.path.join(d, "w.txt")}.get(x, x) for x in v if x != "cvt"]
        mine = subprocess.run([CLI] + args + ["-lm", "1", "-qc-only", "-o", "mine%d" % k, "-outdir", os.path.join(d, "output")],
                              capture_output=True, text=True)
        try:
            out_ref = REF.run_cli(args + ["-lm", "1", "-o", "ref%d" % k], d)
        except RuntimeError as e:
            # the reference refuses this input (a .fam file has one phenotype column: -n 2; -nind against covariate rows, see
            # trim_individuals in gemma_cli.cpp): same refusal, same message
            assert mine.returncode != 0, (v, str(e)[-400:])
            assert _first_error(str(e)) == _first_error(mine.stdout + mine.stderr), (v, str(e)[-400:], mine.stdout[-400:])
            n_refused += 1
            continue
        assert mine.returncode == 0, (v, mine.stdout + mine.stderr)
        for key in COUNT_KEYS:
            assert _count(out_ref, key) == _count(mine.stdout, key), (v, key)
        ref_rows = [ln.split("\t") for ln in open(os.path.join(d, "output", "ref%d.assoc.txt" % k)).read().splitlines()[1:]]
        kept = [m for m in (ln.rstrip("\n").split("\t") for ln in open(os.path.join(d, "output", "mine%d.qc.txt" % k))) if m[1] == "1"]
        assert [r[1] for r in ref_rows] == [m[0] for m in kept], v                                 # same SNPs, same order
        assert [[r[0], r[2], r[5], r[6]] for r in ref_rows] == [m[4:8] for m in kept], v           # chr, ps, allele1, allele0 (annotation / .bim)
        assert [r[3] for r in ref_rows] == [m[2] for m in kept], v                                 # n_miss
        assert [r[7] for r in ref_rows] == ["%.3f" % float(m[3]) for m in kept], v               # af as the reference prints it
        assert len(kept) >= (10 if "-nind" not in v else 1)                                                                     # the case is not degenerate
    assert n_refused >= 1                                  # the refusal branch is exercised (-nind 1000 with unusable individuals)
