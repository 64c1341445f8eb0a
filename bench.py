#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric: SNPs/sec of `-lmm 4` (Wald + LRT + score) at n = 50 000
individuals, SNP-sharded across N B200s (one process per GPU), with the reference CPU path timed
beside it.

A "step" is one pass of the hot path over one batch of synthetic SNPs per GPU:
  PLINK 2-bit genotype rows -> decode + mean-impute -> eigen-projection U^T x (tensor cores)
  -> fused per-SNP lambda search + Wald / LRT / score tests -> SUMSTAT rows.
`value`  : SNPs/s with the step's .bed bytes already resident in HBM (device pointers in/out);
`e2e`    : the same through the host-buffer C-ABI call gb200_lmm_batch_bed (pinned host .bed rows
           in, SUMSTAT rows back to the host inside the timed region);
`roofline`: the dominant kernel (the projection GEMM): algorithmic 2 n^2 flop per SNP over its own
           CUDA-event time, against the measured bf16 peak of MEASURED_PEAKS.json;
`cpu_baseline`: the oracle port of the reference path (OpenBLAS dgemm on all host cores for
           U^T X + the reference's single-threaded per-SNP loop) on a bounded sample.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (no GPU kernels of ours)
"""
import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=50000, help="analysed individuals")
    ap.add_argument("--batch", type=int, default=8192, help="SNPs per step per GPU")
    ap.add_argument("--mode", type=int, default=4, help="-lmm mode (1 Wald, 2 LRT, 3 score, 4 all)")
    ap.add_argument("--utx-path", type=int, default=0, help="0 auto, 1 FP64 tiled, 2 int8 tensor core")
    ap.add_argument("--slices", type=int, default=0, help="int8 planes of U (0 = default 6)")
    ap.add_argument("--cta-pair", type=int, default=-1, help="tensor-core kernels as CTA pairs (cta_group::2): -1 library default, 0, 1")
    ap.add_argument("--overlap", type=int, default=-1, help="sub-batch pipeline (projection || per-SNP tests): -1 default, 0, 1")
    ap.add_argument("--workload", default="lmm", choices=["lmm", "gk"], help="lmm: SNPs/s of -lmm (headline); gk: K=XX^T TFLOP/s")
    ap.add_argument("--lmm-hoist", type=int, default=-1, help="lockstep kernel: hoisted common-lambda passes: -1 default, 0, 1")
    ap.add_argument("--cvt", type=int, default=1, help="covariates incl. intercept (headline config: 1); extra columns are synthetic N(0,1)")
    ap.add_argument("--gk-miss", type=float, default=0.0, help="--workload gk: fraction of missing genotypes in the synthetic data")
    ap.add_argument("--lmm-kernel", type=int, default=0, help="0 auto, 1 warp-per-SNP, 2 lockstep-CTA pipeline")
    ap.add_argument("--cpu-sample", type=int, default=0, help="SNPs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gk", action="store_true", help="skip the -gk K=XX^T TFLOP/s side measurement (rank 0, n=10000)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    bf16_burst=d.get("bf16_tflops", 1590.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


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
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- host-side problem pieces shared by both arms -------------------------------------------------
def host_orthogonal(n, seed, block=2048):
    """Dense orthogonal U (n x n): random orthogonal diagonal blocks mixed by one Householder reflection."""
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


def cpu_reference_sample(n, U, ev, UtW, Uty, n_snps, mode, l_mle_null, logl_mle_H0, snp_offset=0):
    """The reference path on the host: U^T X with OpenBLAS (all cores, like fast_dgemm -> cblas_dgemm) and the
    reference's single-threaded per-SNP loop (oracle port).  Returns (snps_per_s, t_utx, t_opt)."""
    from gemma_b200 import synth
    g = synth.genotypes(n, n_snps, seed=SEED, snp_offset=snp_offset).astype(np.float64)
    X = np.ascontiguousarray(g.T)                                 # n x l, no missing in the perf configs
    t0 = time.perf_counter()
    UtX = U.T @ X
    t1 = time.perf_counter()
    if cpu_kind() == "reference":                                  # the reference's own compiled per-SNP code (oracle/_ref)
        from oracle import ref as REF
        REF.assoc_utx(ev, UtW, Uty, UtX, mode, l_mle_null=l_mle_null, logl_mle_H0=logl_mle_H0)
    else:
        from oracle import oracle as O
        O.lmm_analyze_utx(ev, UtW, Uty, UtX, mode, l_mle_null=l_mle_null, logl_mle_H0=logl_mle_H0)
    t2 = time.perf_counter()
    return n_snps / (t2 - t0), t1 - t0, t2 - t1


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


def cpu_sample_note():
    return ("per-SNP loop = the reference's own src/lmm.cpp compiled against the GSL API shim (oracle/_ref), single-threaded as in the reference"
            if cpu_kind() == "reference" else "per-SNP lambda search = restated oracle port, single-threaded as in the reference")


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the box's host cores: the reference's own per-SNP code compiled
    against the GSL API shim (oracle/_ref) when that library is present, else the restated oracle port; U^T X by OpenBLAS."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gemma_b200 import synth
    from oracle import oracle as O
    n = args.n
    cores = os.cpu_count() or 1
    U = host_orthogonal(n, SEED)
    ev = synth.spectrum_like_kinship(n, SEED)
    gc = synth.genotypes(n, 64, seed=SEED, snp_offset=10 ** 9).astype(np.float64)
    y = synth.phenotype(n, gc, SEED)
    W = np.ones((n, 1))
    UtW = U.T @ W; Uty = U.T @ y
    l_mle, logl = O.calc_lambda_null("L", ev, UtW, Uty)
    sample = args.cpu_sample or max(8, min(64, int(2.0e6 / n)))
    for _ in range(args.warmup):
        cpu_reference_sample(n, U, ev, UtW, Uty, max(2, sample // 8), args.mode, l_mle, logl)
    t0 = time.perf_counter()
    tu = to = 0.0
    for k in range(args.steps):
        _, a, b = cpu_reference_sample(n, U, ev, UtW, Uty, sample, args.mode, l_mle, logl, snp_offset=k * sample)
        tu += a; to += b
    dt = time.perf_counter() - t0
    val = args.steps * sample / dt
    line = {"impl": "reference", "metric": "snps_per_sec_lmm%d" % args.mode, "value": val, "unit": "SNPs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "-lmm %d, n=%d individuals, %d SNPs per step (bounded sample of the per-GPU batch)"
                                   % (args.mode, n, sample), "n": n, "snps_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "SNPs/s", "cores": cores, "kind": cpu_kind(),
                             "sample": "%d SNPs/step x %d steps; U^T X by OpenBLAS dgemm on %d threads (%.1f%% of time), %s (%.1f%%)"
                                       % (sample, args.steps, cores, 100 * tu / (tu + to), cpu_sample_note(), 100 * to / (tu + to))},
            "e2e": {"value": val, "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import gemma_b200
    from gemma_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, B, K, Wm = args.n, args.batch, args.steps, args.warmup
    if Wm < 3:
        Wm = 3
    bps = (n + 3) // 4
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream shared by torch, NCCL and the library
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = gemma_b200.Context(local, stream=stream.cuda_stream)
    ctx.set_option("utx_path", args.utx_path)
    ctx.set_option("n_slices", args.slices)
    if args.lmm_hoist >= 0:
        ctx.set_option("lmm_hoist", args.lmm_hoist)
    ctx.set_option("lmm_kernel", args.lmm_kernel)
    if args.cta_pair >= 0:
        ctx.set_option("cta_pair", args.cta_pair)
    if args.overlap >= 0:
        ctx.set_option("overlap", args.overlap)

    # ---- run-constant state, generated on the device (identical on every rank) ----------------
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    U = torch.zeros((n, n), dtype=torch.float64, device=dev)
    blk = 2048
    for s in range(0, n, blk):
        e = min(n, s + blk)
        q, _ = torch.linalg.qr(torch.randn((e - s, e - s), dtype=torch.float64, device=dev, generator=g))
        U[s:e, s:e] = q
    v = torch.randn(n, dtype=torch.float64, device=dev, generator=g); v /= v.norm()
    w = v @ U
    for s in range(0, n, 4096):
        e = min(n, s + 4096)
        U[s:e] -= 2.0 * v[s:e, None] * w[None, :]
    del q, w
    ev_h = synth.spectrum_like_kinship(n, SEED)
    ev = torch.from_numpy(ev_h).to(dev)
    gc = torch.from_numpy(synth.genotypes(n, 64, seed=SEED, snp_offset=10 ** 9).astype(np.float64))
    y_h = synth.phenotype(n, gc.numpy(), SEED)
    y = torch.from_numpy(y_h).to(dev)
    Wt = torch.ones((args.cvt, n), dtype=torch.float64, device=dev)
    if args.cvt > 1:
        Wt[:args.cvt - 1] = torch.randn((args.cvt - 1, n), dtype=torch.float64, device=dev, generator=g)   # intercept stays LAST
    UtWt = (Wt @ U).contiguous()                                                       # (U^T W)^T, c x n
    Uty = (y @ U).contiguous()
    ctx.lmm_setup_rotated_dev(n, args.cvt, U.data_ptr(), ev.data_ptr(), UtWt.data_ptr(), Uty.data_ptr())
    nm = ctx.lmm_null(float(ev_h.mean()))
    ctx.lmm_params(args.mode, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])

    # ---- per-step inputs: a different SNP batch every step (inputs >> L2; no reuse between steps) ----
    n_batches = K + Wm
    beds = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=(rank * n_batches + k) * B) for k in range(n_batches)]
    out_dev = torch.empty((K, B, 8), dtype=torch.float64, device=dev)
    scratch = torch.empty((B, 8), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_dev(k, dst):
        ctx.lmm_batch_bed_dev(beds[k].data_ptr(), None, n, B, bps, dst.data_ptr())

    for k in range(Wm):
        step_dev(k, scratch)
    torch.cuda.synchronize()
    gathered = torch.empty((world, K, B, 8), dtype=torch.float64, device=dev) if world > 1 else None

    # ---- timed region: K steps + the single gather of the SUMSTAT rows -------------------------
    ctx.profile_enable(True); ctx.profile_reset()
    l0 = ctx.profile_get("__launches")[1]
    sampler = ClockSampler(local); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        step_dev(Wm + k, out_dev[k])
    if world > 1:
        dist.all_gather_into_tensor(gathered, out_dev)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop()
    launches = ctx.profile_get("__launches")[1] - l0 + (1 if world > 1 else 0)
    prof = {k: ctx.profile_get(k) for k in ("utx", "lmm", "decode", "fix")}
    ctx.profile_enable(False)
    value = world * K * B / (ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI call (H2D of the .bed rows + D2H of the SUMSTAT rows) ----
    e2e = None
    if not args.no_e2e:
        hb = [beds[Wm + k].cpu().pin_memory() for k in range(K)]
        hb_np = [b.numpy() for b in hb]
        for k in range(min(2, K)):
            ctx.lmm_batch_bed(hb_np[k], n)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for k in range(K):
            ctx.lmm_batch_bed(hb_np[k], n)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * K * B / (float(ms2.item()) * 1e-3), "unit": "SNPs/s",
               "h2d_bytes_per_step": int(B * bps), "d2h_bytes_per_step": int(B * 64)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = measured_peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")        # per-launch DRAM bytes from the committed ncu capture of this exact config
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("n") == n and tj.get("batch") == B and tj.get("slices") == ctx.get_option("n_slices"):
            traffic = {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in tj.items() if isinstance(v, dict) and "dram_read_bytes" in v}
    utx_ms, utx_n = prof["utx"]
    lmm_ms, lmm_n = prof["lmm"]
    flops_per_launch = 2.0 * n * n * B                    # 2 n^2 per SNP x SNPs per launch
    roof = None
    if utx_n:
        ach = flops_per_launch / (utx_ms / utx_n * 1e-3) / 1e12
        roof = {"kernel": "i8_gemm_kernel (U^T X projection)" if (args.utx_path != 1 and n >= 1024) else "dgemm_kernel (FP64 U^T X)",
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": ach / peaks["bf16"],
                "traffic": traffic.get("i8_gemm_pair_kernel") if (args.utx_path != 1 and n >= 1024 and args.cta_pair != 0) else None,
                "peak_source": peaks["source"] + ", bf16 sustained",
                "note": "algorithmic FP64-equivalent flops 2*n^2 per SNP; the int8 path executes n_slices x as many "
                        "integer MACs (see DESIGN.md)",
                "share_of_step": utx_ms / ms, "avg_launch_ms": utx_ms / utx_n}
        if args.utx_path != 1 and n >= 1024:
            T = ctx.get_option("n_slices")
            roof["executed"] = {"tops_int8": ach * T, "digit_planes": T, "frac_of_2x_bf16_peak": ach * T / (2.0 * peaks["bf16"]),
                                "note": "integer MACs actually issued on the tensor pipe (T exact int8 digit planes of U per "
                                        "FP64-equivalent product); the dense int8 rate of sm_100a is 2x the bf16 rate"}
    lmm_roof = None
    if lmm_n:
        by = (8.0 * n + 64.0) * B
        a = by / (lmm_ms / lmm_n * 1e-3) / 1e9
        lmm_roof = {"kernel": "lmm_assoc_kernel (fused per-SNP tests)", "bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": a / peaks["hbm_gbs"], "share_of_step": lmm_ms / ms, "avg_launch_ms": lmm_ms / lmm_n,
                    "traffic": traffic.get("lmm_assoc_v2_kernel") if args.lmm_kernel != 1 else None,
                    "note": "algorithmic bytes 8n+64 per SNP; the kernel is FP64-pipe bound by construction (~16 lockstep passes, ~1000 FP64 "
                            "instructions per individual and SNP): see DESIGN.md 4.1"}

    # (before the CPU baseline: the host BLAS worker threads it starts keep spinning for a while and slow the synchronous
    #  per-chunk calls of the kinship entry point by an order of magnitude)
    gk = None
    if not args.no_gk and world == 1:               # side measurements: N = 1 only
        try:
            del beds, out_dev, scratch
            torch.cuda.empty_cache()
            g = measure_gk(10000, 65536, 3, 3, ctx=ctx, stream=stream, local=local)
            gk = {"value": g["value"], "unit": g["unit"], "config": g["config"]["workload"], "ms_per_step": g["ms_per_step"],
                  "kernel_tflops": g["roofline"]["achieved"], "frac_of_bf16_peak": g["roofline"]["frac"], "clocks": g["clocks"]}
        except Exception as ex:                     # the side measurement must never cost the headline line
            gk = {"error": str(ex)[:200]}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample = args.cpu_sample or max(8, min(64, int(2.0e6 / n)))
        U_h = U.cpu().numpy()
        UtW_h = UtWt.cpu().numpy().T.copy(); Uty_h = Uty.cpu().numpy()
        cores = os.cpu_count() or 1
        cpu_reference_sample(n, U_h, ev_h, UtW_h, Uty_h, max(2, sample // 8), args.mode, nm["l_mle_null"], nm["logl_mle_H0"])
        v_cpu, tu, to = cpu_reference_sample(n, U_h, ev_h, UtW_h, Uty_h, sample, args.mode, nm["l_mle_null"], nm["logl_mle_H0"])
        cpu = {"value": v_cpu, "unit": "SNPs/s", "cores": cores, "kind": cpu_kind(),
               "sample": "%d SNPs of the same workload; U^T X by OpenBLAS dgemm on %d threads (%.2f s), %s (%.2f s)"
                         % (sample, cores, tu, cpu_sample_note(), to)}

    line = {"metric": "snps_per_sec_lmm%d" % args.mode, "value": value, "unit": "SNPs/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "-lmm %d (Wald+LRT+score), n=%d individuals, %d SNPs per step per GPU from PLINK 2-bit rows, "
                                   "c=%d covariate(s), precomputed eigendecomposition (BASELINE config 4: 5M SNPs sharded by SNP)"
                                   % (args.mode, n, B, args.cvt),
                       "n": n, "snps_per_step_per_gpu": B, "parallelism": "snp-shard x%d, 1 NCCL all-gather of SUMSTAT rows" % world,
                       "l2": "every step reads a different %.0f MB .bed batch and streams %.1f GB of U planes (inputs >> L2)"
                             % (B * bps / 1e6, ctx.get_option("n_slices") * n * n / 1e9)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roof, "roofline_lmm": lmm_roof, "cpu_baseline": cpu, "gk": gk,
            "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in prof.items()}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_gk(n, B, K, Wm, ctx=None, stream=None, local=0, cta_pair=-1, miss=0.0):
    """BASELINE config 2: -gk 1 (centred kinship) on synthetic n x p PLINK genotypes, 1 GPU.
    Reported with the algorithmic flops of ONE triangle, n(n+1)p."""
    import torch
    import gemma_b200
    from gemma_b200 import synth
    dev = torch.device("cuda", local)
    if ctx is None:
        stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
        ctx = gemma_b200.Context(local, stream=stream.cuda_stream)
    if cta_pair >= 0:
        ctx.set_option("cta_pair", cta_pair)
    bps = (n + 3) // 4
    beds = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=k * B, miss_rate=miss) for k in range(K + Wm)]
    ctx.kin_begin(n, 1)
    for k in range(Wm):
        ctx.kin_add_bed_dev(beds[k].data_ptr(), B, bps)
    ctx.kin_finish_dev()
    torch.cuda.synchronize()
    ctx.profile_enable(True); ctx.profile_reset()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.kin_begin(n, 1)
    torch.cuda.synchronize()
    e0.record(stream)
    for k in range(K):
        ctx.kin_add_bed_dev(beds[Wm + k].data_ptr(), B, bps)
    ctx.kin_finish_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    kin_ms, kin_n = ctx.profile_get("kin"); dec_ms, _ = ctx.profile_get("decode"); fix_ms, fix_n = ctx.profile_get("fix")
    p = K * B
    flops = float(n) * (n + 1) * p
    peaks = measured_peaks()
    line = {"metric": "gk_centered_kinship_tflops", "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s (n(n+1)p, one triangle)",
            "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 x int8 -> int32 exact (+FP64 rank-one centring)", "data": "synthetic",
            "config": {"workload": "-gk 1 centred kinship, n=%d individuals, %d SNPs per step (PLINK 2-bit, %s)"
                                   % (n, B, ("%.3g%% missing genotypes" % (100 * miss)) if miss else "no missing"),
                       "n": n, "snps_per_step": B, "snps_per_sec": p / (ms * 1e-3), "missing_rate": miss,
                       "sparse_missing_terms_ms": fix_ms, "sparse_missing_launches": fix_n},
            "clocks": clocks,
            "roofline": {"kernel": "i8_gemm_kernel (mode 1: K += Z Z^T)", "bound": "tensor",
                         "achieved": (flops / (kin_ms * 1e-3) / 1e12) if kin_n else None, "peak": peaks["bf16"], "unit": "TFLOP/s",
                         "frac": (flops / (kin_ms * 1e-3) / 1e12 / peaks["bf16"]) if kin_n else None, "traffic": None,
                         "peak_source": peaks["source"] + ", bf16 sustained", "launches": kin_n, "kernel_ms": kin_ms, "decode_ms": dec_ms,
                         "note": "exact int8 MACs: the int8 tensor rate is 2x the bf16 rate, so frac may reach 2.0 against the bf16 denominator"}}
    ctx.profile_enable(False)
    return line


def run_gk(args):
    """--workload gk [--gpus N]: every rank accumulates K over its own SNP range, one NCCL all-reduce combines them."""
    import torch
    import torch.distributed as dist
    import gemma_b200
    from gemma_b200 import synth, shard
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n if args.n != 50000 else 10000
    if world == 1:
        print(json.dumps(measure_gk(n, max(args.batch, 16384), args.steps, max(3, args.warmup), cta_pair=args.cta_pair, miss=args.gk_miss)))
        return
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    ctx = gemma_b200.Context(local, stream=stream.cuda_stream)
    B, K, Wm = max(args.batch, 16384), args.steps, max(3, args.warmup)
    bps = (n + 3) // 4
    beds = [synth.make_bed_torch(n, B, dev, seed=SEED, snp_offset=(rank * (K + Wm) + k) * B) for k in range(K + Wm)]
    def run(lo, hi):
        ctx.kin_begin(n, 1)
        for k in range(lo, hi):
            ctx.kin_add_bed_dev(beds[k].data_ptr(), B, bps)
        ptr, ns = ctx.kin_finish_dev()
        Kt = shard.device_tensor(ptr, (n, n))
        return shard.combine_partial_kinship(Kt, ns)
    run(0, Wm)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    Kt, ns = run(Wm, Wm + K)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        p = world * K * B
        flops = float(n) * (n + 1) * p
        print(json.dumps({"metric": "gk_centered_kinship_tflops", "value": flops / (ms.item() * 1e-3) / 1e12,
                          "unit": "TFLOP/s (n(n+1)p, one triangle)", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms.item() / K,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
                          "dtype": "int8 x int8 -> int32 exact (+FP64 rank-one centring)",
                          "config": {"workload": "-gk 1, n=%d, %d SNPs per step per GPU, SNP ranges per rank + one NCCL all-reduce of K" % (n, B),
                                     "ns_total": ns, "trace_over_n": float(torch.diagonal(Kt).mean().item())}}))
    dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "gk":
        run_gk(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
