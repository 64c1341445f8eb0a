"""Host side of the GEMMA-compatible CLI (gemma_b200/host/gemma_cli.cpp): readers + SNP QC must select
exactly the SNPs / individuals the reference selects (bit-exact ids and counts).  CPU only (-qc-only)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import refpipe as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gemma_b200", "host")])


def _qc(args, outdir, name):
    _build()
    r = subprocess.run([CLI] + args + ["-qc-only", "-o", name, "-outdir", str(outdir)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(outdir, name + ".qc.txt"))]
    return r.stdout, rows


def test_mouse_qc_matches_reference_selection(golden_dir, tmp_path):
    d = os.path.join(golden_dir, "mouse_hs1940")
    out, rows = _qc(["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt",
                     "-a", d + "/mouse_hs1940.anno.txt"], tmp_path, "mouse")
    assert "## number of analyzed individuals = 1410" in out and "## number of total individuals = 1940" in out
    assert "## number of analyzed SNPs         =    10768" in out and "## number of total SNPs/var        =    12226" in out
    bb = R.Bimbam(d + "/mouse_hs1940.geno.txt.gz")
    ph, ind = R.read_pheno(d + "/mouse_hs1940.pheno.txt", (1,))
    idv, W = R.process_cvt_phen(ind)
    isnp, n_miss, maf = R.qc_bimbam(bb, idv)
    assert [r[0] for r in rows] == bb.rs
    assert np.array_equal(np.array([int(r[1]) for r in rows]), isnp)
    assert np.array_equal(np.array([int(r[2]) for r in rows]), n_miss)
    assert np.allclose(np.array([float(r[3]) for r in rows]), maf, rtol=0, atol=1e-15)


def test_bxd_covariates_maf_r2_qc(golden_dir, tmp_path):
    d = os.path.join(golden_dir, "BXD")
    out, rows = _qc(["-g", d + "/BXD_geno.txt.gz", "-p", d + "/BXD_pheno.txt", "-c", d + "/BXD_covariates2.txt",
                     "-a", d + "/BXD_snps.txt", "-maf", "0.1"], tmp_path, "bxd")
    assert "## number of covariates = 3" in out and "## number of analyzed individuals = 67" in out
    bb = R.Bimbam(d + "/BXD_geno.txt.gz")
    ph, ind = R.read_pheno(d + "/BXD_pheno.txt", (1,))
    rws, icvt = R.read_cvt(d + "/BXD_covariates2.txt")
    idv, W = R.process_cvt_phen(ind, rws, icvt)
    isnp, n_miss, maf = R.qc_bimbam(bb, idv, W, maf_level=0.1)
    assert np.array_equal(np.array([int(r[1]) for r in rows]), isnp)
    assert int(isnp.sum()) == 7317


def test_mouse_nind_snps_selection(golden_dir, tmp_path):
    """-nind 400 -snps list (the LOCO test's inputs, test/dev_tests.rb:57-63): same individuals / SNPs as the oracle."""
    d = os.path.join(golden_dir, "mouse_hs1940")
    out, rows = _qc(["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt",
                     "-a", d + "/mouse_hs1940.anno.txt", "-snps", d + "/mouse_hs1940_snps.txt", "-nind", "400",
                     "-loco", "1"], tmp_path, "loco")
    bb = R.Bimbam(d + "/mouse_hs1940.geno.txt.gz")
    ph, ind = R.read_pheno(d + "/mouse_hs1940.pheno.txt", (1,))
    idv, W = R.process_cvt_phen(ind)
    idv = R.trim_individuals(idv, 400)
    snps = {ln.split()[0] for ln in open(d + "/mouse_hs1940_snps.txt") if ln.strip()}
    isnp, n_miss, maf = R.qc_bimbam(bb, idv, snps=snps)
    assert "## number of total individuals = 400" in out
    assert "## number of analyzed individuals = %d" % int(idv.sum()) in out
    assert np.array_equal(np.array([int(r[1]) for r in rows]), isnp)


def test_cli_rejects_unknown_flags_and_missing_inputs():
    _build()
    r = subprocess.run([CLI, "-bogus"], capture_output=True, text=True)
    assert r.returncode != 0 and "unrecognized option" in r.stdout
    r = subprocess.run([CLI, "-g", "x", "-p", "y", "-lmm", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "missing relatedness file" in r.stdout
    # flag combinations the reference refuses (src/gemma.cpp:1125-1131, src/param.cpp:923-933) or that are outside this framework
    for argv, msg in ((["-g", "x", "-p", "y", "-gk", "-lmm", "1"], "only one of"),
                      (["-g", "x", "-p", "y", "-k", "k", "-lmm", "7"], "not supported"),
                      (["-g", "x", "-p", "y", "-a", "a", "-lm", "1", "-loco", "1"], "LOCO only works with LMM and K"),
                      (["-g", "x", "-p", "y", "-a", "a", "-k", "k", "-lmm", "1", "-loco", "1", "-gxe", "e"], "LOCO does not support GXE"),
                      (["-g", "x", "-p", "y", "-k", "k", "-n", "1", "2", "3", "-lmm", "1"], "two phenotypes"),
                      (["-g", "x", "-p", "y", "-k", "k", "-n", "1", "2", "-lmm", "1", "-gxe", "e"], "multivariate G x E"),
                      (["-p", "y", "-gk"], "need -g and -p")):
        r = subprocess.run([CLI] + argv, capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stdout + r.stderr, (argv, r.stdout, r.stderr)


def test_cli_loco_needs_annotation_and_bimbam_input(golden_dir):
    _build()
    d = os.path.join(golden_dir, "mouse_hs1940")
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt"]
    r = subprocess.run([CLI] + base + ["-gk", "-loco", "1", "-qc-only"], capture_output=True, text=True)
    assert r.returncode != 0 and "LOCO requires annotation file" in r.stdout + r.stderr       # src/param.cpp:924-926
    r = subprocess.run([CLI] + base + ["-a", d + "/mouse_hs1940.anno.txt", "-ksnps", d + "/mouse_hs1940_snps.txt", "-gk", "-loco", "1", "-qc-only"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "LOCO does not allow -ksnps" in r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/example/mouse_hs1940.bed"), reason="reference examples absent")
def test_plink_qc_counts_match_bimbam_counts(tmp_path):
    out, rows = _qc(["-bfile", "/root/reference/example/mouse_hs1940"], tmp_path, "plink")
    # same cohort through the PLINK reader: .fam phenotype column 1, 2-bit genotypes
    assert "## number of total individuals = 1940" in out and "## number of total SNPs/var        =    12226" in out


def test_plink_reader_refuses_malformed_bed(tmp_path):
    """A synthetic PLINK triple (no reference files needed): the valid file passes QC; a truncated payload, a wrong magic number
    and an individual-major file fail loudly instead of decoding garbage."""
    from gemma_b200 import synth
    _build()
    n, l = 37, 25
    bed, G = synth.make_bed(n, l, seed=5, miss_rate=0.02)
    base = str(tmp_path / "toy")
    with open(base + ".fam", "w") as f:
        for i in range(n):
            f.write("f%d i%d 0 0 1 %.3f\n" % (i, i, 0.1 * i))
    with open(base + ".bim", "w") as f:
        for s in range(l):
            f.write("1\trs%d\t0\t%d\tA\tG\n" % (s, 100 + s))
    payload = bytes(bytearray(np.ascontiguousarray(bed).reshape(-1)))
    assert len(payload) == l * ((n + 3) // 4)

    def run(blob):
        with open(base + ".bed", "wb") as f:
            f.write(blob)
        return subprocess.run([CLI, "-bfile", base, "-gk", "-qc-only", "-o", "toy", "-outdir", str(tmp_path)], capture_output=True, text=True)

    r = run(b"\x6c\x1b\x01" + payload)
    assert r.returncode == 0 and "## number of total SNPs/var        =" in r.stdout, r.stdout + r.stderr
    r = run(b"\x6c\x1b\x01" + payload[:-5])
    assert r.returncode != 0 and "truncated .bed file" in r.stdout + r.stderr
    r = run(b"\x00\x00\x01" + payload)
    assert r.returncode != 0 and "not a PLINK .bed file" in r.stdout + r.stderr
    r = run(b"\x6c\x1b\x00" + payload)
    assert r.returncode != 0 and "individual-major" in r.stdout + r.stderr


def test_blank_lines_in_the_genotype_file_do_not_shift_snps(golden_dir, tmp_path):
    """Lines without a token (empty, or blanks / tabs only) are dropped by the line pipeline itself, so the QC pass and the later
    passes (which index the QC flags by line number) number the SNP lines identically."""
    import gzip
    d = os.path.join(golden_dir, "BXD")
    args = ["-p", d + "/BXD_pheno.txt", "-a", d + "/BXD_snps.txt", "-maf", "0.1"]
    _, rows0 = _qc(["-g", d + "/BXD_geno.txt.gz"] + args, tmp_path, "plain")
    lines = gzip.open(d + "/BXD_geno.txt.gz", "rt").read().split("\n")
    for pos, junk in ((3, ""), (40, " \t "), (41, ""), (2000, "\t")):
        lines.insert(pos, junk)
    holes = str(tmp_path / "holes.txt")
    with open(holes, "w") as f:
        f.write("\n".join(lines) + "\n\n")
    _, rows1 = _qc(["-g", holes] + args, tmp_path, "holes")
    assert rows1 == rows0
