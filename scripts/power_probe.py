#!/usr/bin/env python
"""Kernel-isolated power / clock probe (1 GPU): loops ONE stage of the -lmm step for a few seconds each and samples nvidia-smi --
(a) the tensor-core projection alone, (b) the per-SNP tests alone, (c) the whole step -- to tell which stage pulls the SM clock
under the 1 kW cap and what overlapping them could buy.  Writes gpurun_out/power_probe.json.
  python scripts/power_probe.py [--n 50000] [--batch 8192] [--seconds 4] [--opt name=value ...]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gemma_b200
from gemma_b200 import synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--mode", type=int, default=4)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "power_probe.json"))
a = ap.parse_args()
n, B = a.n, a.batch
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
g = torch.Generator(device=dev); g.manual_seed(1)
U, _ = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device=dev, generator=g))
U = U.contiguous()
ev_h = synth.spectrum_like_kinship(n, 1)
y = (U @ torch.from_numpy(synth.polygenic_rotated(ev_h, 1)).to(dev)).contiguous()
ev = torch.from_numpy(ev_h).to(dev)
UtWt = (torch.ones((1, n), dtype=torch.float64, device=dev) @ U).contiguous(); Uty = (y @ U).contiguous()
ctx = gemma_b200.Context(0, stream=stream.cuda_stream)
for kv in a.opt:
    k, v = kv.split("="); ctx.set_option(k, int(v))
ctx.lmm_setup_rotated_dev(n, 1, U.data_ptr(), ev.data_ptr(), UtWt.data_ptr(), Uty.data_ptr())
nm = ctx.lmm_null(float(ev_h.mean()))
ctx.lmm_params(a.mode, l_mle_null=nm["l_mle_null"], logl_mle_H0=nm["logl_mle_H0"])
bps = (n + 3) // 4
beds = [synth.make_bed_torch(n, B, dev, seed=3, snp_offset=k * B) for k in range(3)]
out = torch.empty((B, 8), dtype=torch.float64, device=dev)
res = {"n": n, "snps_per_launch": B, "options": a.opt}
for name, mask in (("projection_only", 1), ("tests_only", 2), ("whole_step", 3)):
    ctx.set_option("stage_mask", 3)
    ctx.lmm_batch_bed_dev(beds[0].data_ptr(), None, n, B, bps, out.data_ptr())      # a valid projection for the tests-only loop
    ctx.set_option("stage_mask", mask)
    torch.cuda.synchronize()
    s = bench.ClockSampler(0); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter(); it = 0
    while time.perf_counter() - t0 < a.seconds:
        ctx.lmm_batch_bed_dev(beds[it % 3].data_ptr(), None, n, B, bps, out.data_ptr()); it += 1
        if it % 4 == 0:
            torch.cuda.synchronize()
    e1.record(stream); torch.cuda.synchronize()
    c = s.stop()
    res[name] = {"ms_per_launch": e0.elapsed_time(e1) / it, "launches": it, **c}
ctx.set_option("stage_mask", 3)
try:
    res["fp64_fma_peak_tflops"] = ctx.measure_fp64_fma(1.0)[0]
except Exception as ex:
    res["fp64_fma_peak_tflops"] = repr(ex)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps(res))
