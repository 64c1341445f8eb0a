"""-widv (residual weights, src/gemma.cpp:2594-2644) in the CLI: the host side of the route -- which individuals stay, the centred and
weighted kinship matrix handed to the eigensolver, the weights that rescale the rows of U -- is taken from `gemma-b200 -qc-only`
(production functions `process_cvt_phen`, `read_kin`, `weighted_kinship`) and carried through the oracle's restatement of the rest of
the -lmm chain; the result has to reproduce the assoc file of the reference's own CLI run with the same -widv file.  What this CPU
test does not execute is the device part of the route (`gb200_eigh` with centring off and the ordinary per-SNP path), both of which
the GPU suite covers on their own."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as REF
from oracle import refpipe as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")


def _read_bin(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"GB2MAT01"
        r, c = struct.unpack("<QQ", f.read(16))
        return np.frombuffer(f.read(), dtype=np.float64).reshape(r, c)


def test_widv_route_reproduces_the_reference_cli(tmp_path):
    if not REF.available():
        pytest.skip("compiled reference not available")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gemma_b200", "host")])
    d = str(tmp_path)
    rng = np.random.default_rng(12)
    n, l = 90, 160
    f = rng.uniform(0.1, 0.5, l)
    G = rng.binomial(2, f[:, None], size=(l, n)).astype(float)
    miss = rng.random((l, n)) < 0.01
    with gzip.open(os.path.join(d, "g.txt.gz"), "wt") as fo:
        for s in range(l):
            fo.write("rs%d, A, G, " % s + ", ".join("NA" if miss[s, i] else "%g" % G[s, i] for i in range(n)) + "\n")
    y = 0.5 * G[3] + 0.4 * G[17] + rng.normal(size=n)
    with open(os.path.join(d, "p.txt"), "w") as fo:
        for i in range(n):
            fo.write(("NA" if i % 13 == 5 else "%.6f" % y[i]) + "\n")
    w = rng.uniform(0.3, 2.5, n)
    with open(os.path.join(d, "w.txt"), "w") as fo:
        for i in range(n):
            fo.write(("NA" if i % 17 == 3 else "%.4f" % w[i]) + "\n")
    base = ["-g", os.path.join(d, "g.txt.gz"), "-p", os.path.join(d, "p.txt")]
    REF.run_cli(base + ["-gk", "-o", "k"], d)
    kfile = os.path.join(d, "output", "k.cXX.txt")
    out_ref = REF.run_cli(base + ["-k", kfile, "-widv", os.path.join(d, "w.txt"), "-lmm", "1", "-o", "refw"], d)
    ref_rows = [ln.split("\t") for ln in open(os.path.join(d, "output", "refw.assoc.txt")).read().splitlines()[1:]]
    # the same run without weights differs visibly (the test would not notice a dropped -widv otherwise)
    REF.run_cli(base + ["-k", kfile, "-lmm", "1", "-o", "ref0"], d)
    ref0 = [ln.split("\t") for ln in open(os.path.join(d, "output", "ref0.assoc.txt")).read().splitlines()[1:]]
    assert len(ref0) != len(ref_rows) or any(abs(float(a[7]) - float(b[7])) > 1e-3 * abs(float(b[7])) for a, b in zip(ref_rows, ref0))

    r = subprocess.run([CLI] + base + ["-k", kfile, "-widv", os.path.join(d, "w.txt"), "-lmm", "1", "-qc-only", "-o", "mine", "-outdir",
                                       os.path.join(d, "output")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for key in ("number of analyzed individuals", "number of analyzed SNPs"):
        a = [ln for ln in out_ref.splitlines() if key in ln][0].split("=")[-1].strip()
        b = [ln for ln in r.stdout.splitlines() if key in ln][0].split("=")[-1].strip()
        assert a == b, key
    Gw = _read_bin(os.path.join(d, "output", "mine.wkin.txt.bin"))
    wv = _read_bin(os.path.join(d, "output", "mine.widv.txt.bin"))[:, 0]
    qc = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(d, "output", "mine.qc.txt"))]
    isnp = np.array([int(q[1]) for q in qc])
    keep = np.array([(i % 13 != 5) and (i % 17 != 3) for i in range(n)])
    assert Gw.shape == (int(keep.sum()),) * 2 and np.allclose(wv, np.round(w, 4)[keep])
    assert [q[0] for q in qc if q[1] == "1"] == [rr[1] for rr in ref_rows]
    # the rest of the chain, restated: eigendecomposition without centring, rows of U scaled, null model, per-SNP Wald test
    U, ev, trace_G = R.eigen_decomp_zeroed(np.ascontiguousarray(Gw))
    U = U * np.sqrt(wv)[:, None]
    W = np.ones((int(keep.sum()), 1)); yk = np.array([float("%.6f" % v) for v in y])[keep]
    UtW, Uty = U.T @ W, U.T @ yk
    l_mle, logl = O.calc_lambda_null("L", ev, UtW, Uty)
    X = O.lmm_impute(np.where(miss, np.nan, G)[np.ix_(np.nonzero(isnp)[0], np.nonzero(keep)[0])])
    got = O.lmm_analyze_utx(ev, UtW, Uty, U.T @ X, 1, l_mle_null=l_mle, logl_mle_H0=logl)
    ref = np.array([[float(x) for x in rr[7:]] for rr in ref_rows])              # beta se logl_H1 l_remle p_wald
    for j, k in ((0, "beta"), (1, "se"), (4, "p_wald")):
        assert np.allclose(got[k], ref[:, j], rtol=3e-6, atol=0), k              # 7 printed digits
    assert np.allclose(got["lambda_remle"], ref[:, 3], rtol=1e-4, atol=2e-5)
