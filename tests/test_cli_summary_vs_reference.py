"""The command lines of the CLI's GPU tests (mouse -gk / -lmm, the LOCO + -nind + -snps run of test/dev_tests.rb:57-77, the two-phenotype
run) with `-qc-only`: everything the host does before the first kernel -- flag handling, readers, individual / SNP selection, the
kinship file written by the reference read back with -nind in force -- must give the summary block the reference's own CLI prints
("## number of ..." lines of PARAM::PrintSummary, src/param.cpp:1100-1140), line for line.  CPU only."""
import os
import subprocess

import pytest

from oracle import ref as REF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")


def _summary(txt):
    return [ln.strip() for ln in txt.splitlines() if ln.startswith("## number")]


def test_summary_block_matches_the_reference_cli(golden_dir, tmp_path):
    if not REF.available():
        pytest.skip("compiled reference not available")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gemma_b200", "host")])
    cwd = str(tmp_path)
    out_dir = os.path.join(cwd, "output")
    d = os.path.join(golden_dir, "mouse_hs1940")
    base = ["-g", d + "/mouse_hs1940.geno.txt.gz", "-p", d + "/mouse_hs1940.pheno.txt", "-a", d + "/mouse_hs1940.anno.txt"]
    loco = base + ["-snps", d + "/mouse_hs1940_snps.txt", "-nind", "400", "-loco", "1"]
    k, kl = os.path.join(out_dir, "rk.cXX.txt"), os.path.join(out_dir, "rlk.cXX.txt")
    cases = [("gk", base + ["-gk"]),
             ("lmm", base + ["-n", "1", "-k", k, "-lmm"]),
             ("loco gk", loco + ["-gk"]),
             ("loco lmm", loco + ["-n", "1", "-k", kl, "-lmm", "-no-check"]),
             ("two phenotypes", base + ["-n", "1", "6", "-k", k, "-lmm"])]
    names = {"gk": "rk", "loco gk": "rlk"}
    for tag, args in cases:
        ref_out = REF.run_cli(args + ["-o", names.get(tag, "r")], cwd)
        r = subprocess.run([CLI] + args + ["-qc-only", "-o", "m", "-outdir", out_dir], capture_output=True, text=True)
        assert r.returncode == 0, (tag, r.stdout + r.stderr)
        want, got = _summary(ref_out), _summary(r.stdout)
        assert len(want) >= 6 and got == want, (tag, want, got)
