"""bench.py's reference arm runs on host cores only, so its contract (one JSON line, the keys the driver reads, rank 0 alone under
torch.distributed.run) can be checked without a GPU at a toy size."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--individuals", "300", "--steps", "2", "--warmup", "3", "--cpu-sample", "8"]


def _check_line(out, gpus):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "snps_per_sec_lmm4" and d["unit"] == "SNPs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == gpus and d["steps"] == 2 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "SNPs/step" in cb["sample"]
    assert d["config"]["n"] == 300 and "workload" in d["config"] and "model" not in d["config"]
    return d


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 1)


def test_reference_arm_under_torchrun_only_rank0_works():
    """Launched like the driver does for N > 1: rank 0 prints the line with all host threads (torch.distributed.run exports
    OMP_NUM_THREADS=1, which the arm overrides), the other rank exits 0 without output."""
    import socket
    with socket.socket() as sk:                                     # a port nobody holds right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2"] + ARGS
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_line(r.stdout, 2)
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


@pytest.mark.parametrize("workload,metric", [("gk", "gk_centered_kinship_tflops"), ("lmm1", "snps_per_sec_lmm1")])
def test_reference_arm_other_workloads(workload, metric):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--n", "256",
                        "--steps", "1", "--warmup", "3", "--cpu-sample", "256" if workload == "gk" else "8"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == metric and d["value"] > 0
