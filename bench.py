#!/usr/bin/env python
"""bench.py -- the BASELINE.json metrics of the GEMMA -gk / -eigen / -lmm hot path on N B200s (one process per GPU), with the
reference CPU path timed beside them.

Workloads (--workload):
  lmm  (default, BASELINE config 4's per-GPU shard, the headline): SNPs/sec of `-lmm 4` (Wald + LRT + score) at n = 50 000.
       Setup, untimed but reported: K from synthetic PLINK genotypes on the int8 tensor path (gb200_kin_*), its
       eigendecomposition on the device (gb200_eigh_dev, wall time = `eigh_s`), null model.  A "step" is one pass of the hot path
       over one batch of synthetic SNPs per GPU:
         PLINK 2-bit rows -> decode + mean-impute -> eigen-projection U^T x (tensor cores) -> fused per-SNP lambda search +
         Wald / LRT / score tests -> SUMSTAT rows;   SNPs shard across ranks, one NCCL all-gather of the SUMSTAT rows at the end.
  lmm1 (BASELINE config 3): `-lmm 1` (Wald), n = 10 000.
  gk   (BASELINE config 2): `-gk 1` centred kinship, n = 10 000 x 500 000 SNPs per step; N > 1: SNP ranges + one all-reduce of K.
  mv   (BASELINE config 5): multivariate LMM, two phenotypes, n = 10 000.

`value`   : whole-job throughput with the step's .bed bytes already resident in HBM (device pointers in / out);
`e2e`     : the same through the host-buffer C-ABI call (pinned host .bed rows in, result rows back to the host, copies inside
            the timed region);
`roofline`: the dominant kernel of the workload, algorithmic work over its own CUDA-event time, against MEASURED_PEAKS.json;
`parity`  : rows produced INSIDE the timed region (and a batch with 1 % missing genotypes) against the reference's own per-SNP
            code (oracle/_ref) on the same inputs: max relative deviation per column (north star: 1e-6);
`cpu_baseline`: the reference CPU path on this box's host cores on a bounded sample of the same workload.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (none of our kernels)
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260923
K_SNP_OFFSET = 2 * 10 ** 9       # the kinship SNPs: a range disjoint from the tested SNPs (i.i.d. generator: any disjoint subset is "every 10th SNP")

DEFAULTS = {   # workload -> (n, SNPs per step per GPU)
    "lmm": (50000, 65536), "lmm1": (10000, 262144), "gk": (10000, 500000), "mv": (10000, 32768)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lmm", choices=["lmm", "lmm4", "lmm1", "gk", "mv"])
    ap.add_argument("--n", "--individuals", dest="n", type=int, default=0,
                    help="analysed individuals (0 = the workload's BASELINE size); spell it --individuals under torch.distributed.run, whose own parser trips over --n")
    ap.add_argument("--batch", type=int, default=0, help="SNPs per step per GPU (0 = the workload's default)")
    ap.add_argument("--mode", type=int, default=0, help="-lmm mode (0 = the workload's: 4 for lmm, 1 for lmm1 / mv)")
    ap.add_argument("--cvt", type=int, default=1, help="covariates incl. intercept (BASELINE configs: 1); extra columns are synthetic N(0,1)")
    ap.add_argument("--miss", type=float, default=0.0, help="fraction of missing genotypes in the timed batches (BASELINE perf configs: 0)")
    ap.add_argument("--u-source", default="eigh", choices=["eigh", "qr"],
                    help="eigh: U, eval = gb200_eigh_dev of the kinship matrix built by gb200_kin_* (default); qr: Haar U + synthetic spectrum (fast setup for kernel A/B runs)")
    ap.add_argument("--k-snps", type=int, default=0, help="SNPs in the kinship matrix of the setup (0 = 10 x n)")
    ap.add_argument("--utx-path", type=int, default=0, help="0 auto, 1 FP64 tiled, 2 int8 tensor core")
    ap.add_argument("--slices", type=int, default=0, help="int8 digit planes of U (0 = library default)")
    ap.add_argument("--cta-pair", type=int, default=-1)
    ap.add_argument("--overlap", type=int, default=-1)
    ap.add_argument("--lmm-hoist", type=int, default=-1)
    ap.add_argument("--lmm-kernel", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="SNPs per internal sub-batch of the bed entry points (0 = auto)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (A/B runs)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="SNPs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gk", action="store_true", help="skip the short -gk side measurement of the lmm line")
    a = ap.parse_args()
    if a.workload == "lmm4":
        a.workload = "lmm"
    n0, b0 = DEFAULTS[a.workload]
    a.n = a.n or n0
    a.batch = a.batch or b0
    a.mode = a.mode or (4 if a.workload == "lmm" else 1)
    a.warmup = max(3, a.warmup)
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    bf16_burst=d.get("bf16_tflops", 1590.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


def committed_traffic(kernel, n, snps_per_launch):
    """Per-launch DRAM bytes of `kernel` from the committed ncu capture of this n (profiles/r02_traffic.json, written by
    scripts/ncu_extract.py), scaled by SNPs per launch when the run's launches differ from the captured 8192-SNP launch (both
    kernels stream per SNP; the capture itself is one launch)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    if not t or t.get("n") != n:
        return None
    return (t["dram_read_bytes"] + t["dram_write_bytes"]) * float(snps_per_launch) / float(t["snps_per_launch"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads(n):
    """Give the host BLAS all the cores (torch.distributed.run exports OMP_NUM_THREADS=1, which made round 1's reference arm
    run its dgemm on one thread for N > 1)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(n); os.environ["OPENBLAS_NUM_THREADS"] = str(n)


def workload_text(args):
    n, B = args.n, args.batch
    if args.workload == "lmm":
        return ("-lmm %d (Wald+LRT+score), n=%d individuals, %d SNPs per step per GPU from PLINK 2-bit rows, c=%d covariate(s); "
                "U, eval from the device eigendecomposition of the synthetic kinship matrix (BASELINE config 4: 5M SNPs sharded by SNP)"
                % (args.mode, n, B, args.cvt))
    if args.workload == "lmm1":
        return "-lmm %d (Wald), n=%d individuals, %d SNPs per step per GPU from PLINK 2-bit rows, c=%d, eigendecomposition supplied (BASELINE config 3)" % (args.mode, n, B, args.cvt)
    if args.workload == "gk":
        return "-gk 1 centred kinship, n=%d individuals x %d SNPs per step per GPU (PLINK 2-bit), SNP ranges per rank + one all-reduce of K (BASELINE config 2)" % (n, B)
    return "multivariate LMM -lmm %d, 2 phenotypes, n=%d individuals, %d SNPs per step per GPU from PLINK 2-bit rows, c=%d (BASELINE config 5)" % (args.mode, n, B, args.cvt)


def make_config(args, world):
    """Identical in both arms (the reference arm times a bounded sample of THIS configuration)."""
    return {"workload": workload_text(args), "n": args.n, "snps_per_step_per_gpu": args.batch, "missing_rate": args.miss,
            "parallelism": "snp-shard x%d, 1 NCCL %s" % (world, "all-reduce of K" if args.workload == "gk" else "all-gather of the result rows"),
            "l2": "every step streams a different >= 0.6 GB .bed batch (and the U digit planes): inputs >> the 126 MB L2, no flush needed"}


METRIC = {"lmm": ("snps_per_sec_lmm%d", "SNPs/s"), "lmm1": ("snps_per_sec_lmm%d", "SNPs/s"), "mv": ("snps_per_sec_mvlmm%d", "SNPs/s"),
          "gk": ("gk_centered_kinship_tflops", "TFLOP/s (n(n+1)p, one triangle)")}


# ---- CPU side ---------------------------------------------------------------------------------------------------------------
_CPU_KIND = None


def cpu_kind():
    """"reference": src/lmm.cpp itself, compiled in place against the GSL API shim (oracle/_ref/libgemma_ref.so, prebuilt in the
    authoring container and shipped with the snapshot); "port": the restated oracle when that library is not there."""
    global _CPU_KIND
    if _CPU_KIND is None:
        try:
            from oracle import ref as REF
            REF.lib()
            _CPU_KIND = "reference"
        except Exception:
            _CPU_KIND = "port"
    return _CPU_KIND


def cpu_note():
    return ("per-SNP loop = the reference's own src/lmm.cpp compiled against the GSL API shim (oracle/_ref), single-threaded as in the reference"
            if cpu_kind() == "reference" else "per-SNP lambda search = restated oracle port, single-threaded as in the reference")


def cpu_assoc(ev, UtW, Uty, UtX, mode, l_mle_null, logl_mle_H0):
    if cpu_kind() == "reference":
        from oracle import ref as REF
        return REF.assoc_utx(ev, UtW, Uty, UtX, mode, l_mle_null=l_mle_null, logl_mle_H0=logl_mle_H0)
    from oracle import oracle as O
    return O.lmm_analyze_utx(ev, UtW, Uty, UtX, mode, l_mle_null=l_mle_null, logl_mle_H0=logl_mle_H0)


def cpu_lmm_sample(n, U, ev, UtW, Uty, n_snps, mode, l_mle_null, logl_mle_H0, snp_offset=0):
    """The reference path on the host for one sample: U^T X with OpenBLAS (all cores: fast_dgemm -> cblas_dgemm, src/lmm.cpp:1521)
    and the reference's single-threaded per-SNP loop (:1526-1562).  Returns (seconds, t_utx, t_opt, rows)."""
    from gemma_b200 import synth
    g = synth.genotypes(n, n_snps, seed=SEED, snp_offset=snp_offset).astype(np.float64)
    X = np.ascontiguousarray(g.T)                                 # n x l, no missing in the perf configs
    t0 = time.perf_counter()
    UtX = U.T @ X
    t1 = time.perf_counter()
    rows = cpu_assoc(ev, UtW, Uty, UtX, mode, l_mle_null, logl_mle_H0)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1, rows


def cpu_gk_sample(n, n_snps):
    """BimbamKin / PlinkKin's dgemm on one K_BATCH_SIZE-like batch (src/gemma_io.cpp:1554,1711): centred genotypes, K += X X^T."""
    from gemma_b200 import synth
    g = synth.genotypes(n, n_snps, seed=SEED, snp_offset=K_SNP_OFFSET).astype(np.float64)
    g -= g.mean(axis=1, keepdims=True)
    X = np.ascontiguousarray(g.T)                                 # n x l like Xlarge
    t0 = time.perf_counter()
    K = X @ X.T
    dt = time.perf_counter() - t0
    return dt, float(K[0, 0])


def host_orthogonal(n, seed, block=2048):
    """Dense orthogonal U for the CPU arm (the host cannot eigendecompose a 50 000-wide matrix in minutes; the CPU timings do not
    depend on the values of U): random orthogonal diagonal blocks mixed by one Householder reflection."""
    rng = np.random.default_rng(seed)
    U = np.zeros((n, n))
    for s in range(0, n, block):
        e = min(n, s + block)
        q, _ = np.linalg.qr(rng.standard_normal((e - s, e - s)))
        U[s:e, s:e] = q
    v = rng.standard_normal(n); v /= np.linalg.norm(v)
    w = v @ U
    for s in range(0, n, 4096):
        e = min(n, s + 4096)
        U[s:e] -= 2.0 * v[s:e, None] * w[None, :]
    return U


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores, on a bounded sample per step of
    the b200 arm's configuration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    host_threads(cores)
    from gemma_b200 import synth
    from oracle import oracle as O
    n = args.n
    metric, unit = METRIC[args.workload]
    K, Wm = args.steps, max(1, min(args.warmup, 2))
    if args.workload == "gk":
        sample = args.cpu_sample or max(512, min(20000, int(2.0e7 / n)))
        for _ in range(Wm):
            cpu_gk_sample(n, max(256, sample // 8))
        dt = 0.0
        for _ in range(K):
            dt += cpu_gk_sample(n, sample)[0]                      # the dgemm only (not the synthetic-genotype generator)
        val = float(n) * (n + 1) * sample * K / dt / 1e12
        note = "%d SNPs/step x %d steps: the K += X X^T dgemm of BimbamKin/PlinkKin (OpenBLAS, %d threads) on centred FP64 genotypes; flops counted as n(n+1)p like the GPU arm (the dgemm executes 2 n^2 p)" % (sample, K, cores)
        metric_name = metric
    else:
        mode = args.mode
        U = host_orthogonal(n, SEED)
        ev = synth.spectrum_like_kinship(n, SEED)
        W = np.ones((n, 1))
        UtW = U.T @ W; Uty = synth.polygenic_rotated(ev, SEED)              # the same phenotype model as the b200 arm (pve 0.5)
        l_mle, logl = O.calc_lambda_null("L", ev, UtW, Uty)
        if args.workload == "mv":
            # the reference's multivariate per-SNP loop is not part of oracle/_ref's library; its CLI cannot take rotated input.
            print(json.dumps({"impl": "reference", "unavailable": "mvLMM per-SNP loop of the reference is only reachable through its CLI on files; no bounded in-memory sample"}))
            return
        sample = args.cpu_sample or max(16, min(256, int(4.0e6 / n)))
        for _ in range(Wm):
            cpu_lmm_sample(n, U, ev, UtW, Uty, max(4, sample // 8), mode, l_mle, logl)
        tu = to = 0.0
        for k in range(K):
            _, a, b, _ = cpu_lmm_sample(n, U, ev, UtW, Uty, sample, mode, l_mle, logl, snp_offset=k * sample)
            tu += a; to += b
        dt = tu + to                                               # dgemm + per-SNP loop (not the synthetic-genotype generator)
        val = K * sample / dt
        note = ("%d SNPs/step x %d steps of the same workload (the reference's batches hold 20000 SNP columns; its dgemm is more efficient there: "
                "the U^T X share below is an upper bound); U^T X by OpenBLAS dgemm on %d threads (%.1f%% of the time), %s (%.1f%%)"
                % (sample, K, cores, 100 * tu / (tu + to), cpu_note(), 100 * to / (tu + to)))
        metric_name = metric % mode
    line = {"impl": "reference", "metric": metric_name, "value": val, "unit": unit,
            "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": 1e3 * dt / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args, args.gpus),
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": cpu_kind(), "sample": note},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- GPU side ---------------------------------------------------------------------------------------------------------------
class Env:
    """torch / NCCL plumbing shared by the workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import gemma_b200
        self.torch, self.dist, self.gb = torch, dist, gemma_b200
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(minutes=60))
        self.stream = torch.cuda.Stream(device=self.dev)       # a real (non-default) stream shared by torch, NCCL and the library
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0

    def context(self, args):
        ctx = self.gb.Context(self.local, stream=self.stream.cuda_stream)
        ctx.set_option("utx_path", args.utx_path)
        ctx.set_option("n_slices", args.slices)
        ctx.set_option("lmm_kernel", args.lmm_kernel)
        if args.lmm_hoist >= 0:
            ctx.set_option("lmm_hoist", args.lmm_hoist)
        if args.cta_pair >= 0:
            ctx.set_option("cta_pair", args.cta_pair)
        if args.overlap >= 0:
            ctx.set_option("overlap", args.overlap)
        if args.chunk:
            ctx.set_option("batch_chunk", args.chunk)
        for kv in args.opt:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        return ctx

    def wait_for_rank0(self, key):
        """Ranks > 0 wait on the rendezvous store (not inside a pending NCCL collective) while rank 0 runs the minutes-long setup."""
        if self.world == 1:
            return
        store = self.dist.distributed_c10d._get_default_store()
        if self.rank == 0:
            store.set(key, "1")
        else:
            store.wait([key], datetime.timedelta(hours=2))

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_scalars(self, vals):
        """list of floats per rank -> [world][len] on every rank"""
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [list(map(float, t.tolist()))]
        out = self.torch.empty((self.world, len(vals)), dtype=self.torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        return out.tolist()

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_eigensystem(env, args, n):
    """Run-constant state of a -lmm run, produced by the library itself: K = XX^T/p on the int8 tensor path from synthetic PLINK rows
    (gb200_kin_*), centred and eigendecomposed on the device (gb200_eigh_dev).  Rank 0 computes, one NCCL broadcast distributes
    (SURVEY 8e).  Returns (U [n,n] device tensor, ev device tensor, info dict)."""
    torch, dist = env.torch, env.dist
    from gemma_b200 import synth
    dev = env.dev
    info = {"u_source": args.u_source}
    U = torch.empty((n, n), dtype=torch.float64, device=dev)
    ev = torch.empty(n, dtype=torch.float64, device=dev)
    tr = torch.zeros(1, dtype=torch.float64, device=dev)
    if env.rank == 0:
        if args.u_source == "qr":
            g = torch.Generator(device=dev); g.manual_seed(SEED)
            t0 = time.perf_counter()
            A = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
            Q, _ = torch.linalg.qr(A)
            U.copy_(Q); del A, Q
            ev_h = synth.spectrum_like_kinship(n, SEED)
            ev.copy_(torch.from_numpy(ev_h)); tr[0] = float(ev_h.mean())
            torch.cuda.synchronize()
            info["qr_s"] = time.perf_counter() - t0
        else:
            ctxk = env.gb.Context(env.local, stream=env.stream.cuda_stream)
            ks = args.k_snps or 10 * n
            bps = (n + 3) // 4
            chunk = max(1024, min(32768, (1 << 29) // bps // 128 * 128))
            t0 = time.perf_counter()
            ctxk.kin_begin(n, 1)
            done = 0
            while done < ks:
                lc = min(chunk, ks - done)
                bed = synth.make_bed_torch(n, lc, dev, seed=SEED, snp_offset=K_SNP_OFFSET + done)
                ctxk.kin_add_bed_dev(bed.data_ptr(), lc, bps)
                torch.cuda.synchronize()
                del bed
                done += lc
            kptr, ns = ctxk.kin_finish_dev()
            torch.cuda.synchronize()
            info["kinship_s"] = time.perf_counter() - t0            # includes generating the synthetic rows on the device
            info["kinship_snps"] = int(ns)
            t0 = time.perf_counter()
            trace_G, n_zero, n_neg = ctxk.eigh_dev(kptr, n, U.data_ptr(), ev.data_ptr(), center=True)
            torch.cuda.synchronize()
            info["eigh_s"] = time.perf_counter() - t0
            info["eigh_workspace_gb"] = ctxk.get_option("eigh_workspace_bytes") / 1e9
            info["eigh_zero_eigenvalues"] = int(n_zero)
            tr[0] = trace_G
            ctxk.close()
            torch.cuda.empty_cache()
    env.wait_for_rank0("eigensystem")
    if env.world > 1:
        t0 = time.perf_counter()
        dist.broadcast(U, src=0); dist.broadcast(ev, src=0); dist.broadcast(tr, src=0)
        torch.cuda.synchronize()
        info["broadcast_s"] = time.perf_counter() - t0
    info["trace_G"] = float(tr.item())
    return U, ev, info


def parity_check(ctx, n, mode, U_h, ev_h, UtW_h, Uty_h, nm, got_rows, G_rows, label):
    """GPU rows vs the reference's own per-SNP code (oracle/_ref; the oracle port when that library is absent) on the same SNPs."""
    from oracle import oracle as O
    X = O.lmm_impute(np.where(G_rows < 0, np.nan, G_rows.astype(np.float64)))        # n x l
    UtX = U_h.T @ X
    ref = cpu_assoc(ev_h, UtW_h, Uty_h, UtX, mode, nm["l_mle_null"], nm["logl_mle_H0"])
    cols = {1: ("beta", "se", "p_wald"), 2: ("p_lrt",), 3: ("beta", "se", "p_score"), 4: ("beta", "se", "p_wald", "p_lrt", "p_score")}[mode]
    out = {"snps": int(len(ref)), "against": "oracle/_ref (the reference's src/lmm.cpp)" if cpu_kind() == "reference" else "oracle port", "what": label}
    worst = 0.0
    for k in cols + ("lambda_remle", "lambda_mle", "logl_H1"):
        a, b = np.asarray(got_rows[k], float), np.asarray(ref[k], float)
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            out[k] = "NaN pattern differs"; worst = float("inf"); continue
        m = ~np.isnan(b)
        e = float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300))) if m.any() else 0.0
        out[k] = e
        if k in cols:
            worst = max(worst, e)
    out["max_rel_err"] = worst
    out["within_1e-6"] = bool(worst < 1e-6)
    return out


def run_lmm(args):
    env = Env(args)
    torch, dist = env.torch, env.dist
    from gemma_b200 import synth, SUMSTAT_DTYPE
    world, rank, dev, stream = env.world, env.rank, env.dev, env.stream
    n, B, K, Wm, mode = args.n, args.batch, args.steps, args.warmup, args.mode
    bps = (n + 3) // 4
    ctx = env.context(args)

    # ---- run-constant state ---------------------------------------------------------------------------------
    t_setup = time.perf_counter()
    U, ev, setup = build_eigensystem(env, args, n)
    ev_h = ev.cpu().numpy()
    # phenotype drawn from the model itself (pve 0.5): y = U (sqrt(h2 ev / mean ev + 1 - h2) * z), so that every SNP's REML / ML
    # root is interior and the per-SNP kernel runs its full grid + Brent + Newton search
    y = (U @ torch.from_numpy(synth.polygenic_rotated(ev_h, SEED)).to(dev)).contiguous()
    g = torch.Generator(device=dev); g.manual_seed(SEED + 1)
    Wt = torch.ones((args.cvt, n), dtype=torch.float64, device=dev)
    if args.cvt > 1:
        Wt[:args.cvt - 1] = torch.randn((args.cvt - 1, n), dtype=torch.float64, device=dev, generator=g)   # intercept stays LAST
    UtWt = (Wt @ U).contiguous()                                                       # (U^T W)^T, c x n   (CalcUtX)
    Uty = (y @ U).contiguous()
    ctx.lmm_setup_rotated_dev(n, args.cvt, U.data_ptr(), ev.data_ptr(), UtWt.data_ptr(), Uty.data_ptr())
    nm = ctx.lmm_null(setup["trace_G"])
    ctx.lmm_params(mode, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
    setup["null_model"] = {k: nm[k] for k in ("l_mle_null", "l_remle_null", "pve_null")}

    # ---- per-step inputs: a cycle of distinct batches, each far larger than L2 ----------------------------------
    nb = min(K + Wm, 4)
    off = lambda k: (rank * nb + k) * B
    beds = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=off(k), miss_rate=args.miss) for k in range(nb)]
    out_dev = torch.empty((K, B, 8), dtype=torch.float64, device=dev)
    scratch = torch.empty((B, 8), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    setup["setup_s"] = time.perf_counter() - t_setup

    def step_dev(k, dst):
        ctx.lmm_batch_bed_dev(beds[k % nb].data_ptr(), None, n, B, bps, dst.data_ptr())

    for k in range(Wm):
        step_dev(k, scratch)
    torch.cuda.synchronize()
    gathered = torch.empty((world, K, B, 8), dtype=torch.float64, device=dev) if world > 1 else None

    # ---- timed region: K steps + the single gather of the SUMSTAT rows -------------------------
    ctx.lmm_counters(reset=True)
    ctx.profile_enable(True); ctx.profile_reset()
    l0 = ctx.profile_get("__launches")[1]
    sampler = ClockSampler(env.local); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    for k in range(K):
        step_dev(Wm + k, out_dev[k])
    e2.record(stream)
    if world > 1:
        dist.all_gather_into_tensor(gathered, out_dev)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_local, ms_compute = e0.elapsed_time(e1), e0.elapsed_time(e2)
    ms = env.max_over_ranks(ms_local)
    clocks = sampler.stop()
    launches = ctx.profile_get("__launches")[1] - l0 + (1 if world > 1 else 0)
    prof = {k: ctx.profile_get(k) for k in ("utx", "lmm", "decode", "fix")}
    counters = ctx.lmm_counters()
    ctx.profile_enable(False)
    value = world * K * B / (ms * 1e-3)
    per_rank = env.gather_scalars([ms_local, ms_compute, prof["utx"][0], prof["lmm"][0], clocks["sm_mhz"] or 0.0, clocks["sm_mhz_min"] or 0.0,
                                   clocks["power_w"] or 0.0])

    # ---- e2e: host buffers through the C-ABI call (H2D of the .bed rows + D2H of the SUMSTAT rows inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        nh = min(K, 3)
        hb = [beds[(Wm + k) % nb].cpu().pin_memory() for k in range(nh)]
        hb_np = [b.numpy() for b in hb]
        ctx.lmm_batch_bed(hb_np[0], n)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for k in range(K):
            ctx.lmm_batch_bed(hb_np[k % nh], n)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = env.max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * K * B / (ms2 * 1e-3), "unit": "SNPs/s", "h2d_bytes_per_step": int(B * bps), "d2h_bytes_per_step": int(B * 64),
               "ms_per_step": ms2 / K, "note": "gb200_lmm_batch_bed: pinned host rows, double-buffered H2D on a copy stream under the kernels"}
        del hb, hb_np

    # ---- -gk at N > 1 (BASELINE metric: "-gk K = XX^T TFLOP/s at 1/2/4/8 B200"): every rank accumulates K over its own SNP range, one
    # all-reduce combines them; a short side measurement inside the default line so that the driver's scaling runs carry it
    gk_multi = None
    if not args.no_gk and world > 1 and args.workload == "lmm":
        gk_multi = measure_gk_sharded(env, 10000, 131072, 4, 2)

    if rank != 0:
        env.finish()
        return

    # ---- parity of rows from the timed region, and of a batch with missing genotypes --------------------------------------
    parity = None
    if not args.no_parity and world == 1:
        try:
            host_threads(os.cpu_count() or 1)
            U_h = U.cpu().numpy(); UtW_h = UtWt.cpu().numpy().T.copy(); Uty_h = Uty.cpu().numpy()
            PL = 32
            rows = out_dev[0, :PL].cpu().numpy().reshape(-1).view(SUMSTAT_DTYPE)
            G0 = synth.genotypes(n, PL, seed=SEED, snp_offset=off(Wm % nb), miss_rate=args.miss)
            parity = {"timed_rows": parity_check(ctx, n, mode, U_h, ev_h, UtW_h, Uty_h, nm, rows, G0,
                                                 "first %d SNPs of the first timed step (missing rate %g)" % (PL, args.miss))}
            bm = synth.make_bed_torch(n, 256, dev, seed=SEED, snp_offset=7 * 10 ** 8, miss_rate=0.01)
            om = torch.empty((256, 8), dtype=torch.float64, device=dev)
            ctx.lmm_batch_bed_dev(bm.data_ptr(), None, n, 256, bps, om.data_ptr())
            torch.cuda.synchronize()
            Gm = synth.genotypes(n, PL, seed=SEED, snp_offset=7 * 10 ** 8, miss_rate=0.01)
            parity["missing_1pct"] = parity_check(ctx, n, mode, U_h, ev_h, UtW_h, Uty_h, nm, om[:PL].cpu().numpy().reshape(-1).view(SUMSTAT_DTYPE), Gm,
                                                  "%d SNPs of an extra batch with 1%% missing genotypes (mean-imputed: src/lmm.cpp:1819-1827)" % PL)
            parity["max_rel_err"] = max(parity["timed_rows"]["max_rel_err"], parity["missing_1pct"]["max_rel_err"])
            parity["digit_planes"] = ctx.get_option("n_slices")
        except Exception as ex:                      # the check must never cost the line
            parity = {"error": repr(ex)[:300]}

    # ---- rooflines -------------------------------------------------------------------------------------------------
    peaks = measured_peaks()
    T = ctx.get_option("n_slices")
    chunk = ctx.get_option("batch_chunk")
    i8 = (args.utx_path != 1 and n >= 1024)
    utx_ms, utx_n = prof["utx"]
    lmm_ms, lmm_n = prof["lmm"]
    roof = lmm_roof = None
    if utx_n:
        snps_per_launch = K * B / utx_n
        ach = 2.0 * n * n * snps_per_launch / (utx_ms / utx_n * 1e-3) / 1e12            # 2 n^2 flop per SNP
        roof = {"kernel": "i8_gemm_pair_kernel (U^T X projection, tcgen05 int8)" if i8 else "dgemm_kernel (FP64 U^T X)",
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": ach / peaks["bf16"],
                "traffic": committed_traffic("i8_gemm_pair_kernel", n, snps_per_launch) if i8 else None,
                "algorithmic_bytes_per_launch": T * n * n + snps_per_launch * (n + 8.0 * n) if i8 else None,
                "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
                "note": "algorithmic FP64-equivalent flops 2*n^2 per SNP; the int8 path executes n_slices x as many integer MACs",
                "share_of_step": utx_ms / ms_compute, "avg_launch_ms": utx_ms / utx_n, "snps_per_launch": snps_per_launch, "launches": utx_n}
        if i8:
            roof["executed"] = {"tops_int8": ach * T, "digit_planes": T, "frac_of_2x_bf16_peak": ach * T / (2.0 * peaks["bf16"]),
                                "note": "integer MACs actually issued on the tensor pipe; the dense int8 rate of sm_100a is 2x the bf16 rate"}
    if lmm_n:
        snps_per_launch = K * B / lmm_n
        by = (8.0 * n + 64.0) * snps_per_launch
        a = by / (lmm_ms / lmm_n * 1e-3) / 1e9
        lmm_roof = {"kernel": "lmm_assoc_v2_kernel (fused per-SNP tests)", "bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": a / peaks["hbm_gbs"], "share_of_step": lmm_ms / ms_compute, "avg_launch_ms": lmm_ms / lmm_n,
                    "snps_per_launch": snps_per_launch, "traffic": committed_traffic("lmm_assoc_v2_kernel", n, snps_per_launch),
                    "note": "algorithmic bytes 8n+64 per SNP (SURVEY 8d); the kernel is FP64-issue bound by construction, see `fp64`"}
        try:
            fp64_peak, _ = ctx.measure_fp64_fma(0.5)
            nc = args.cvt
            nidx = (nc + 3) * (nc + 2) // 2
            n_c = (n + 511) // 512 * 512
            # executed FP64 flops (FMA = 2) per individual: hoisted slot = 1 mul + 2(c+2) FMA, + (c+2) products per pass of 5 slots;
            # refinement pass (2 lambdas): nidx products + 2 x (den FMA, reciprocal = 4 FMA, (P-1) mul, P x (1 add + nidx FMA)), P = powers
            f_common = counters["common_slots"] * (1 + 4.0 * (nc + 2)) + (counters["common_slots"] / 5.0) * (nc + 2)
            f_p2 = counters["two_power_passes"] * (nidx + 2 * (2 + 8 + 1 + 2 * (1 + 2 * nidx)))
            f_p3 = counters["three_power_passes"] * (nidx + 2 * (2 + 8 + 2 + 3 * (1 + 2 * nidx)))
            flops = (f_common + f_p2 + f_p3) * n_c
            ach64 = flops / (lmm_ms * 1e-3) / 1e12
            lmm_roof["fp64"] = {"achieved": ach64, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach64 / fp64_peak,
                                "peak_source": "measured here: independent DFMA chains on all SMs for 0.5 s (gb200_measure_fp64_fma)",
                                "passes_per_snp": {"hoisted_lambda_slots": counters["common_slots"] / max(1, counters["snps"]),
                                                   "exact_two_power_passes": counters["two_power_passes"] / max(1, counters["snps"]),
                                                   "exact_three_power_passes": counters["three_power_passes"] / max(1, counters["snps"]),
                                                   "exact_with_logdet": counters["with_logdet"] / max(1, counters["snps"])},
                                "interpolated_refinement": bool(ctx.get_option("lmm_interp")),
                                "note": "executed flops from the kernel's own pass counters (log / special functions not counted)"}
        except Exception as ex:
            lmm_roof["fp64"] = {"error": repr(ex)[:200]}

    # (before the CPU baseline: the host BLAS worker threads it starts keep spinning for a while and slow the synchronous
    #  per-chunk calls of the kinship entry point by an order of magnitude)
    gk = None
    if not args.no_gk and world == 1 and args.workload == "lmm":
        try:
            del beds, out_dev, scratch
            torch.cuda.empty_cache()
            gl = measure_gk(env, 10000, 131072, 4, 3, ctx=ctx)
            gk = {"value": gl["value"], "unit": gl["unit"], "config": gl["config"]["workload"], "ms_per_step": gl["ms_per_step"],
                  "kernel_tflops": gl["roofline"]["achieved"], "frac_of_bf16_peak": gl["roofline"]["frac"], "clocks": gl["clocks"],
                  "note": "short side measurement; BASELINE config 2 proper is `bench.py --workload gk` (profiles/)"}
        except Exception as ex:                     # the side measurement must never cost the headline line
            gk = {"error": str(ex)[:200]}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        host_threads(cores)
        sample = args.cpu_sample or max(16, min(256, int(4.0e6 / n)))
        if parity is None or "error" in parity:
            U_h = U.cpu().numpy(); UtW_h = UtWt.cpu().numpy().T.copy(); Uty_h = Uty.cpu().numpy()
        cpu_lmm_sample(n, U_h, ev_h, UtW_h, Uty_h, max(4, sample // 8), mode, nm["l_mle_null"], nm["logl_mle_H0"])
        dt, tu, to, _ = cpu_lmm_sample(n, U_h, ev_h, UtW_h, Uty_h, sample, mode, nm["l_mle_null"], nm["logl_mle_H0"])
        cpu = {"value": sample / dt, "unit": "SNPs/s", "cores": cores, "kind": cpu_kind(),
               "sample": "%d SNPs of the same workload on the same U; U^T X by OpenBLAS dgemm on %d threads (%.2f s), %s (%.2f s)"
                         % (sample, cores, tu, cpu_note(), to)}

    line = {"metric": METRIC[args.workload][0] % mode, "value": value, "unit": "SNPs/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": make_config(args, world),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "timed_s": ms * 1e-3,
            "roofline": roof, "roofline_lmm": lmm_roof, "parity": parity, "cpu_baseline": cpu, "gk": gk if gk is not None else gk_multi,
            "setup": setup, "eigh_s": setup.get("eigh_s"),
            "digit_planes": T, "internal_sub_batch": chunk,
            "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in prof.items()},
            "per_rank": [{"ms_total": r[0], "ms_compute": r[1], "ms_utx": r[2], "ms_lmm": r[3], "sm_mhz": r[4], "sm_mhz_min": r[5], "power_w": r[6]}
                         for r in per_rank]}
    print(json.dumps(line))
    env.finish()


# ---- -gk -------------------------------------------------------------------------------------------------------------------
def measure_gk(env, n, B, K, Wm, ctx=None, cta_pair=-1, miss=0.0):
    """-gk 1 (centred kinship) on synthetic n x B PLINK genotypes per step, this rank only.  A step = kin_begin .. kin_finish over
    B SNPs; reported with the algorithmic flops of ONE triangle, n(n+1)p."""
    torch = env.torch
    from gemma_b200 import synth
    dev, stream = env.dev, env.stream
    if ctx is None:
        ctx = env.gb.Context(env.local, stream=stream.cuda_stream)
    if cta_pair >= 0:
        ctx.set_option("kin_cta_pair", cta_pair)
    bps = (n + 3) // 4
    sub = min(B, 32768)                                            # rows handed to one gb200_kin_add_bed_dev call
    nb = min(K + Wm, 3)
    beds = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=K_SNP_OFFSET + (env.rank * nb + k) * B, miss_rate=miss) for k in range(nb)]

    def step(k):
        ctx.kin_begin(n, 1)
        b = beds[k % nb]
        for s0 in range(0, B, sub):
            lc = min(sub, B - s0)
            ctx.kin_add_bed_dev(b.data_ptr() + s0 * bps, lc, bps)
        return ctx.kin_finish_dev()

    for k in range(Wm):
        step(k)
    torch.cuda.synchronize()
    ctx.profile_enable(True); ctx.profile_reset()
    l0 = ctx.profile_get("__launches")[1]
    sampler = ClockSampler(env.local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for k in range(K):
        kptr, ns = step(Wm + k)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    kin_ms, kin_n = ctx.profile_get("kin"); dec_ms, _ = ctx.profile_get("decode"); fix_ms, fix_n = ctx.profile_get("fix")
    launches = ctx.profile_get("__launches")[1] - l0
    ctx.profile_enable(False)
    p = K * B
    flops = float(n) * (n + 1) * p
    peaks = measured_peaks()
    return {"metric": "gk_centered_kinship_tflops", "value": flops / (ms * 1e-3) / 1e12, "unit": METRIC["gk"][1],
            "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "ms": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 x int8 -> int32 exact (+FP64 rank-one centring)", "data": "synthetic",
            "config": {"workload": "-gk 1 centred kinship, n=%d individuals, %d SNPs per step (PLINK 2-bit, %s)"
                                   % (n, B, ("%.3g%% missing genotypes" % (100 * miss)) if miss else "no missing")},
            "gpu_launches": int(launches), "clocks": clocks, "kptr": kptr, "ns": ns, "snps_per_sec": p / (ms * 1e-3),
            "sparse_missing_terms_ms": fix_ms,
            "roofline": {"kernel": "i8_gemm_kernel (mode 1: K += Z Z^T, tcgen05 int8)", "bound": "tensor",
                         "achieved": (flops / (kin_ms * 1e-3) / 1e12) if kin_n else None, "peak": peaks["bf16"], "unit": "TFLOP/s",
                         "frac": (flops / (kin_ms * 1e-3) / 1e12 / peaks["bf16"]) if kin_n else None, "traffic": None,
                         "peak_source": peaks["source"] + ", bf16 sustained", "launches": kin_n, "kernel_ms": kin_ms, "decode_ms": dec_ms,
                         "share_of_step": kin_ms / ms if kin_n else None,
                         "note": "exact int8 MACs: the int8 tensor rate is 2x the bf16 rate, so frac may reach 2.0 against the bf16 denominator"}}


def measure_gk_sharded(env, n, B, K, Wm):
    """-gk 1 on `world` GPUs: rank r accumulates K over its own B SNPs per step (gb200_kin_*), shard.combine_partial_kinship rescales
    and all-reduces the n^2 doubles.  Every rank calls this; returns the dict on every rank.  A rank that fails before the first
    collective is reported through a MIN all-reduce of a success flag, so that no rank waits in a collective the others never reach."""
    torch, dist = env.torch, env.dist
    from gemma_b200 import synth, shard
    bps = (n + 3) // 4
    ok, err, ctxg, beds = 1.0, "", None, None

    def local(k):
        ctxg.kin_begin(n, 1)
        for s0 in range(0, B, 32768):
            ctxg.kin_add_bed_dev(beds[k % 2].data_ptr() + s0 * bps, min(32768, B - s0), bps)
        return ctxg.kin_finish_dev()

    try:
        ctxg = env.gb.Context(env.local, stream=env.stream.cuda_stream)
        beds = [synth.make_bed_torch(n, B, env.dev, seed=SEED, snp_offset=K_SNP_OFFSET + (env.rank * 2 + k) * B) for k in range(2)]
        local(0)
        torch.cuda.synchronize()
    except Exception as ex:
        ok, err = 0.0, str(ex)[:200]
    flag = torch.tensor([ok], dtype=torch.float64, device=env.dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if flag.item() < 1.0:
        return {"error": "a rank failed before the first collective: " + err}
    for k in range(Wm):
        ptr, ns = local(k)
        shard.combine_partial_kinship(shard.device_tensor(ptr, (n, n)), ns)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(env.stream)
    for k in range(K):
        ptr, ns = local(Wm + k)
        Kt, ns_tot = shard.combine_partial_kinship(shard.device_tensor(ptr, (n, n)), ns)
    e1.record(env.stream)
    torch.cuda.synchronize()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    tr = float(torch.diagonal(Kt).mean().item())
    ctxg.close()
    return {"value": float(n) * (n + 1) * env.world * K * B / (ms * 1e-3) / 1e12, "unit": METRIC["gk"][1], "n_gpus": env.world,
            "config": "-gk 1 centred kinship, n=%d individuals, %d SNPs per step per GPU (PLINK 2-bit), SNP ranges per rank + one all-reduce of K" % (n, B),
            "ms_per_step": ms / K, "steps": K, "all_reduce_bytes_per_step": int(n * n * 8), "ns_total": ns_tot, "trace_over_n": tr,
            "note": "short side measurement of the default line at N > 1; BASELINE config 2 proper is `bench.py --workload gk --gpus N`"}


def run_gk(args):
    """--workload gk: every rank accumulates K over its own SNP range; one NCCL all-reduce combines them (N > 1)."""
    env = Env(args)
    torch, dist = env.torch, env.dist
    from gemma_b200 import shard
    n, B, K, Wm = args.n, args.batch, args.steps, args.warmup
    bps = (n + 3) // 4
    ctx = env.context(args)
    if args.cta_pair >= 0:
        ctx.set_option("kin_cta_pair", args.cta_pair)
    e2e = None
    if env.world == 1:
        line = measure_gk(env, n, B, K, Wm, ctx=ctx, miss=args.miss)
        ms = line.pop("ms"); line.pop("kptr"); line.pop("ns")
        if not args.no_e2e:
            from gemma_b200 import synth
            hb = synth.make_bed_torch(n, B, env.dev, seed=SEED, snp_offset=K_SNP_OFFSET, miss_rate=args.miss).cpu().pin_memory()
            hb_np = hb.numpy()
            sub = 32768

            Kpin = torch.empty((n, n), dtype=torch.float64).pin_memory()       # K comes back into pinned memory (0.8 GB: pageable D2H runs at a few GB/s)
            Kpin_np = Kpin.numpy()

            def step_host():
                ctx.kin_begin(n, 1)
                for s0 in range(0, B, sub):
                    ctx.kin_add_bed(hb_np[s0:s0 + sub])
                return ctx.kin_finish(out=Kpin_np)

            Kh, _ = step_host()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(2, min(K, 5))
            torch.cuda.synchronize(); e0.record(env.stream)
            for _ in range(reps):
                Kh, _ = step_host()
            e1.record(env.stream); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / reps
            e2e = {"value": float(n) * (n + 1) * B / (ms2 * 1e-3) / 1e12, "unit": METRIC["gk"][1], "h2d_bytes_per_step": int(B * bps),
                   "d2h_bytes_per_step": int(n * n * 8), "ms_per_step": ms2,
                   "note": "gb200_kin_add_bed (host .bed rows) ... gb200_kin_finish (K copied back to the host: 0.8 GB at n = 10 000)"}
    else:
        from gemma_b200 import synth
        nb = min(K + Wm, 3)
        beds = [synth.make_bed_torch(n, B, env.dev, seed=SEED, snp_offset=K_SNP_OFFSET + (env.rank * nb + k) * B, miss_rate=args.miss) for k in range(nb)]
        sub = min(B, 32768)

        def step(k):
            ctx.kin_begin(n, 1)
            for s0 in range(0, B, sub):
                ctx.kin_add_bed_dev(beds[k % nb].data_ptr() + s0 * bps, min(sub, B - s0), bps)
            ptr, ns = ctx.kin_finish_dev()
            return shard.combine_partial_kinship(shard.device_tensor(ptr, (n, n)), ns)

        for k in range(Wm):
            step(k)
        sampler = ClockSampler(env.local); sampler.start()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.profile_get("__launches")[1]
        e0.record(env.stream)
        for k in range(K):
            Kt, ns = step(Wm + k)
        e1.record(env.stream)
        torch.cuda.synchronize()
        ms = env.max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop()
        flops = float(n) * (n + 1) * env.world * K * B
        line = {"metric": "gk_centered_kinship_tflops", "value": flops / (ms * 1e-3) / 1e12, "unit": METRIC["gk"][1], "n_gpus": env.world,
                "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "data": "synthetic", "dtype": "int8 x int8 -> int32 exact (+FP64 rank-one centring)", "clocks": clocks,
                "gpu_launches": int(ctx.profile_get("__launches")[1] - l0 + 2 * K),
                "all_reduce_bytes_per_step": int(n * n * 8), "ns_total": ns, "trace_over_n": float(torch.diagonal(Kt).mean().item())}
    if env.rank == 0:
        line["config"] = make_config(args, env.world)
        line["e2e"] = e2e
        if not args.no_cpu_baseline and env.world == 1:
            cores = os.cpu_count() or 1
            host_threads(cores)
            sample = args.cpu_sample or max(512, min(20000, int(2.0e7 / n)))
            cpu_gk_sample(n, max(256, sample // 8))
            dt, _ = cpu_gk_sample(n, sample)
            line["cpu_baseline"] = {"value": float(n) * (n + 1) * sample / dt / 1e12, "unit": METRIC["gk"][1], "cores": cores, "kind": "reference",
                                    "gflops_executed": 2.0 * n * n * sample / dt / 1e9,
                                    "sample": "%d SNPs (one batch): the K += X X^T cblas_dgemm of PlinkKin (src/gemma_io.cpp:1711) by OpenBLAS on %d threads, "
                                              "centred FP64 genotypes; value counts n(n+1)p like the GPU arm, gflops_executed the 2 n^2 p the dgemm performs" % (sample, cores)}
        print(json.dumps(line))
    env.finish()


# ---- mvLMM (BASELINE config 5) ---------------------------------------------------------------------------------------------
def run_mv(args):
    env = Env(args)
    torch, dist = env.torch, env.dist
    from gemma_b200 import synth
    world, rank, dev, stream = env.world, env.rank, env.dev, env.stream
    n, B, K, Wm, mode = args.n, args.batch, args.steps, args.warmup, args.mode
    bps = (n + 3) // 4
    ctx = env.context(args)
    U, ev, setup = build_eigensystem(env, args, n)
    U_h = U.cpu().numpy(); ev_h = ev.cpu().numpy()
    y1 = U_h @ synth.polygenic_rotated(ev_h, SEED); y2 = 0.4 * y1 + U_h @ synth.polygenic_rotated(ev_h, SEED + 5)   # pve 0.5 each, correlated
    Y = np.stack([y1, y2], axis=1)
    W = np.ones((n, args.cvt))
    if args.cvt > 1:
        W[:, :args.cvt - 1] = np.random.default_rng(SEED).standard_normal((n, args.cvt - 1))
    ctx.mvlmm_setup(U_h, ev_h, W, Y)
    ctx.mvlmm_null()
    nb = min(K + Wm, 3)
    hb = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=(rank * nb + k) * B, miss_rate=args.miss).cpu().pin_memory() for k in range(nb)]
    hb_np = [b.numpy() for b in hb]
    for k in range(Wm):
        ctx.mvlmm_batch_bed(hb_np[k % nb], n, a_mode=mode)
    ctx.profile_enable(True); ctx.profile_reset()
    l0 = ctx.profile_get("__launches")[1]
    sampler = ClockSampler(env.local); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        ctx.mvlmm_batch_bed(hb_np[(Wm + k) % nb], n, a_mode=mode)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    prof = {k: ctx.profile_get(k) for k in ("utx", "lmm", "decode", "fix")}
    launches = ctx.profile_get("__launches")[1] - l0
    if rank == 0:
        peaks = measured_peaks()
        v = world * K * B / (ms * 1e-3)
        lmm_ms, lmm_n = prof["lmm"]
        dev_ms = sum(prof[k][0] for k in prof)          # device work of the steps without the host <-> device copies
        by = (8.0 * n + 64.0) * K * B
        line = {"metric": METRIC["mv"][0] % mode, "value": world * K * B / (dev_ms * 1e-3) if dev_ms > 0 else v,
                "unit": "SNPs/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": make_config(args, world), "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": v, "unit": "SNPs/s", "h2d_bytes_per_step": int(B * bps), "d2h_bytes_per_step": int(B * 64),
                        "note": "gb200_mvlmm_batch_bed with pinned host rows (the only multivariate entry point: `value` is the sum of its kernels' event times)"},
                "roofline": {"kernel": "mv_assoc_kernel (per-SNP EM + Newton-Raphson + tests)", "bound": "hbm",
                             "achieved": by / (lmm_ms * 1e-3) / 1e9 if lmm_n else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": by / (lmm_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if lmm_n else None, "traffic": None,
                             "share_of_step": lmm_ms / ms if lmm_n else None,
                             "note": "algorithmic bytes 8n+64 per SNP; the kernel is FP64 / latency bound (iterative EM over moment sums), DESIGN.md 7"},
                "setup": setup, "kernel_ms": {k: {"ms": x[0], "launches": x[1]} for k, x in prof.items()}}
        print(json.dumps(line))
    env.finish()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "gk":
        run_gk(args)
    elif args.workload == "mv":
        run_mv(args)
    else:
        run_lmm(args)


if __name__ == "__main__":
    main()
