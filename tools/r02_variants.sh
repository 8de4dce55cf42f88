#!/bin/bash
# Bench variants of the persistent CR kernel (round 2 tuning): readiness flags on/off, dealing.
out=gpurun_out
mkdir -p $out
for flags in 1 0; do
  for deal in aware plain; do
    IPXGPU_FUSED_FLAGS=$flags IPXGPU_BAND_DEAL=$deal timeout 200 python bench.py --no-cpu-baseline \
      > $out/var_f${flags}_${deal}.json 2> $out/var_f${flags}_${deal}.err
    python - <<P
import json
try:
    d=json.load(open("$out/var_f${flags}_${deal}.json"))
    r=d["roofline"]
    print("flags=$flags deal=$deal value=%.0f apply_us=%.2f frac=%.3f iso=%.1f s1=%.1f s2=%.1f" % (d["value"], r["apply_us_in_loop"], r["frac"], r["apply_us_isolated_l2_flushed"], r["sweep1_us"], r["sweep2_us"]))
except Exception as e:
    print("flags=$flags deal=$deal FAILED", e)
P
  done
done
IPXGPU_FUSED_FLAGS=1 IPXGPU_FUSED_TRACE=1 timeout 200 python bench.py --no-cpu-baseline --steps 1 --warmup 3 2>&1 >/dev/null | tail -2 | cut -c1-900
IPXGPU_FUSED_FLAGS=0 IPXGPU_FUSED_TRACE=1 timeout 200 python bench.py --no-cpu-baseline --steps 1 --warmup 3 2>&1 >/dev/null | tail -2 | cut -c1-900
