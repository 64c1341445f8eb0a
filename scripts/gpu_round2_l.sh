#!/bin/bash
# GPU session L: A/B of the projection raster / L2 hints (Haar U, 3 steps each; same box, back to back).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="python bench.py --u-source qr --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-gk --no-parity"
for v in "base:" "p6:--opt gemm_panel=6" "p6h:--opt gemm_panel=6 --opt gemm_l2hint=1" "p9h:--opt gemm_l2hint=1" "p12h:--opt gemm_panel=12 --opt gemm_l2hint=1" "base2:"; do
  tag=${v%%:*}; opt=${v#*:}
  timeout 600 $B $opt > gpurun_out/l_ab_$tag.json 2> gpurun_out/l_ab_$tag.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_ab_*.json")):
    try:
        d=json.load(open(f)); k=d["kernel_ms"]
        print(f.split("_ab_")[1], "value %.0f" % d["value"], "utx/launch %.2f ms" % (k["utx"]["ms"]/k["utx"]["launches"]), "lmm/launch %.2f" % (k["lmm"]["ms"]/k["lmm"]["launches"]), d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
    except Exception as e: print(f, "ERR", e)
PY
