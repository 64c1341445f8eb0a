#!/bin/bash
# A/B runs of bench.py on the GPU box; each line: tag + value + per-kernel ms.
# usage: bench_ab.sh "<tag>|<bench args>[|ENV=VAL ...]" ...
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%|*}"; rest="${spec#*|}"; args="${rest%%|*}"; envs=""
  if [[ "$rest" == *"|"* ]]; then envs="${rest#*|}"; fi
  env $envs timeout 900 python bench.py $args > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import sys, json
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % tag).read().strip().splitlines()[-1])
    km = {k: round(v["ms"] / max(1, v["launches"]), 3) for k, v in d.get("kernel_ms", {}).items()}
    if "kernel_ms" not in d: km = d.get("roofline")
    print(tag, "value=%.1f" % d["value"], "e2e=%s" % (d.get("e2e") and round(d["e2e"]["value"])), "ms/step=%.2f" % d["ms_per_step"], km,
          "clk=%s" % d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(tag, "FAILED", e); print(open("gpurun_out/ab_%s.err" % tag).read()[-1500:])
PY
done
