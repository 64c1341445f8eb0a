"""N > 1 host logic on CPU: world_size-2 gloo run of the SNP sharding + single gather."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from gemma_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return str(sk.getsockname()[1])


def test_snp_ranges_partition_in_order():
    for n, w in ((10, 3), (0, 2), (5, 8), (5000000, 8), (17, 1)):
        r = [shard.snp_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_preserves_snp_order(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch.distributed as dist
        from gemma_b200 import shard
        from gemma_b200.api import SUMSTAT_DTYPE
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 11
        lo, hi = shard.snp_range(n, rank, world)
        loc = np.zeros(hi - lo, dtype=SUMSTAT_DTYPE)
        loc["beta"] = np.arange(lo, hi); loc["p_wald"] = np.arange(lo, hi) * 0.5; loc["logl_H1"] = -np.arange(lo, hi)
        out = shard.gather_sumstat(loc, n)
        if rank == 0:
            assert np.array_equal(out["beta"], np.arange(n)) and np.array_equal(out["p_wald"], np.arange(n) * 0.5)
            assert np.array_equal(out["logl_H1"], -np.arange(n)) and len(out) == n
            print("GATHER_OK")
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", _free_port(), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert "GATHER_OK" in r.stdout, r.stdout + r.stderr


def test_two_rank_gloo_partial_kinship_reduce(tmp_path):
    script = tmp_path / "k.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from gemma_b200 import shard
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        rng = np.random.default_rng(5)
        n, p = 9, 31
        X = rng.standard_normal((n, p))
        lo, hi = shard.snp_range(p, rank, world)
        Kloc = torch.from_numpy(X[:, lo:hi] @ X[:, lo:hi].T / (hi - lo))
        K, ns = shard.combine_partial_kinship(Kloc, hi - lo)
        assert ns == p and np.allclose(K.numpy(), X @ X.T / p, atol=1e-13)
        if rank == 0: print("KIN_OK")
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", _free_port(), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert "KIN_OK" in r.stdout, r.stdout + r.stderr
