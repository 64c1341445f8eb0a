"""The CLI's kinship reader (-k with -km 1 / -km 2) against the reference's own ReadFile_kin
(src/gemma_io.cpp:1186-1294, compiled in place: oracle/_ref/libgemma_ref.so) on the same files: individuals without a phenotype
dropped from rows and columns, mixed separators and number formats, a gzip-compressed file, the id-pair list with unknown ids,
pairs in either order and repeated (equal) entries.  `gemma-b200 -qc-only -k ...` writes the matrix it would analyse.  CPU only."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import ref as REF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gemma_b200", "host", "gemma-b200")


def _read_bin(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"GB2MAT01"
        r, c = struct.unpack("<QQ", f.read(16))
        return np.frombuffer(f.read(), dtype=np.float64).reshape(r, c)


def _plink_case(d, n, l, seed):
    from gemma_b200 import synth
    rng = np.random.default_rng(seed)
    bed, _ = synth.make_bed(n, l, seed=seed, miss_rate=0.01)
    base = os.path.join(d, "pl")
    with open(base + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01])); f.write(np.ascontiguousarray(bed).tobytes())
    with open(base + ".bim", "w") as f:
        for s in range(l):
            f.write("1\trs%d\t0\t%d\tA\tG\n" % (s, 100 + s))
    ids = ["id_%03d" % i for i in rng.permutation(n)]
    idv = (rng.random(n) > 0.12).astype(np.int32)
    with open(base + ".fam", "w") as f:
        for i in range(n):
            f.write("fam%d %s 0 0 1 %s\n" % (i, ids[i], ("%.4f" % rng.normal()) if idv[i] else "-9"))
    return base, ids, idv


def _run(args, d, name):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gemma_b200", "host")])
    r = subprocess.run([CLI] + args + ["-qc-only", "-o", name, "-outdir", d], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return _read_bin(os.path.join(d, name + ".kin.txt.bin"))


@pytest.mark.parametrize("seed", [3, 4])
def test_kinship_readers_match_the_reference(tmp_path, seed):
    if not REF.available():
        pytest.skip("compiled reference not available")
    d = str(tmp_path)
    n, l = 57, 40
    rng = np.random.default_rng(100 + seed)
    base, ids, idv = _plink_case(d, n, l, seed)
    A = rng.standard_normal((n, 2 * n)); K = A @ A.T / (2 * n)
    K[3, 5] = K[5, 3] = 0.0                                                  # an exact zero entry
    # ---- -km 1: n x n text, separators and formats mixed (strtok " ,\t" + atof), once plain and once gzipped
    def fmt(v, i, j):
        return ["%.10g", "%.12e", "%g", "%+.9f"][(i + j) % 4] % v
    k1 = os.path.join(d, "k1.txt")
    with open(k1, "w") as f:
        for i in range(n):
            sep = ["\t", " ", ",", ", "][i % 4]
            f.write(sep.join(fmt(K[i, j], i, j) for j in range(n)) + "\n")
    with open(k1, "rb") as f, gzip.open(k1 + ".gz", "wb") as g:
        g.write(f.read())
    want = REF.read_kin(k1, idv, 1)
    assert want.shape == (int(idv.sum()),) * 2
    for path in (k1, k1 + ".gz"):
        got = _run(["-bfile", base, "-k", path, "-km", "1"], d, "km1")
        assert np.array_equal(got, want)                                     # the same atof on the same tokens: bit-identical
    # the BIMBAM route reads the same file with the phenotype file's missingness instead of the .fam's
    with open(os.path.join(d, "ph.txt"), "w") as f:
        for i in range(n):
            f.write("%s\n" % ("0.5" if idv[i] else "NA"))
    with gzip.open(os.path.join(d, "g.txt.gz"), "wt") as f:
        for s in range(l):
            f.write("rs%d, A, G, " % s + ", ".join("%d" % v for v in rng.integers(0, 3, n)) + "\n")
    got = _run(["-g", os.path.join(d, "g.txt.gz"), "-p", os.path.join(d, "ph.txt"), "-k", k1, "-km", "1"], d, "km1b")
    assert np.array_equal(got, want)
    # ---- -km 2: "id1 id2 value" rows keyed by the .fam individual ids
    k2 = os.path.join(d, "k2.txt")
    with open(k2, "w") as f:
        pairs = [(i, j) for i in range(n) for j in range(i + 1)]
        for t, k in enumerate(rng.permutation(len(pairs))):
            i, j = pairs[k]
            if (i + j) % 3 == 0:
                i, j = j, i                                                  # either order
            sep = ["\t", " ", ","][t % 3]
            f.write(sep.join([ids[i], ids[j], "%.10g" % K[i, j]]) + "\n")
            if t % 17 == 0:
                f.write(sep.join([ids[j], ids[i], "%.10g" % K[i, j]]) + "\n")   # repeated, equal: accepted
            if t % 23 == 0:
                f.write("nobody %s 0.25\n" % ids[i])                         # unknown id: skipped
    want2 = REF.read_kin(k2, idv, 2, ids=ids)
    got2 = _run(["-bfile", base, "-k", k2, "-km", "2"], d, "km2")
    assert np.array_equal(got2, want2)
    keep = idv == 1
    assert np.allclose(want2, K[np.ix_(keep, keep)], rtol=1e-9, atol=1e-12)   # and it is the matrix that was written


def test_matrix_and_vector_writers_are_byte_identical_to_the_reference(tmp_path):
    """SURVEY 8 row a5 (WriteMatrix / WriteVector, src/param.cpp:1886-1935).  The reference's CLI writes .cXX.txt / .sXX.txt (-gk 1/2)
    and .eigenU.txt / .eigenD.txt (-eigen) for a small cohort; gemma-b200 reads each file with its production readers and writes it
    back with its production writers: the 10-significant-digit text has to come back byte for byte (reader and formatter both exact)."""
    if not REF.available():
        pytest.skip("compiled reference not available")
    d = str(tmp_path)
    n, l = 41, 150
    base, ids, _ = _plink_case(d, n, l, 9)
    with open(base + ".fam", "w") as f:                                      # every individual has a phenotype: complete matrices
        for i in range(n):
            f.write("fam%d %s 0 0 1 %.4f\n" % (i, ids[i], np.sin(i)))
    for k_mode, suffix in ((1, "cXX"), (2, "sXX")):
        REF.run_cli(["-bfile", base, "-gk", str(k_mode), "-o", "ref"], d)
        ref_file = os.path.join(d, "output", "ref.%s.txt" % suffix)
        _run(["-bfile", base, "-k", ref_file], os.path.join(d, "output"), "mine%d" % k_mode)
        assert open(os.path.join(d, "output", "mine%d.kin.txt" % k_mode), "rb").read() == open(ref_file, "rb").read()
    REF.run_cli(["-bfile", base, "-k", os.path.join(d, "output", "ref.cXX.txt"), "-eigen", "-o", "ref"], d)
    fu, fd = os.path.join(d, "output", "ref.eigenU.txt"), os.path.join(d, "output", "ref.eigenD.txt")
    _run(["-bfile", base, "-k", os.path.join(d, "output", "ref.cXX.txt"), "-u", fu, "-d", fd], os.path.join(d, "output"), "mine_e")
    assert open(os.path.join(d, "output", "mine_e.eigU.txt"), "rb").read() == open(fu, "rb").read()
    assert open(os.path.join(d, "output", "mine_e.eigD.txt"), "rb").read() == open(fd, "rb").read()
    assert os.path.getsize(fu) > n * n * 5 and os.path.getsize(fd) > n * 5
