#!/bin/bash
out=gpurun_out
mkdir -p $out
run() { tag=$1; shift; env "$@" timeout 300 python tools/solve_lp.py random:100000:1000000:10 --impl gpu --crossover 0 --stop-at-switch -1 --per-iter --out $out/r02g_c2_$tag.json > $out/r02g_c2_$tag.log 2>&1; echo "$tag rc=$?"; python - <<P
import json
d=json.load(open("$out/r02g_c2_$tag.json"))["results"]["gpu"]
print("$tag", d["iter"], d["kktiter1"], d["pobjval"], [(r["kktiter"], r["mu"], r["pres"]) for r in d["per_iter"][:8]])
P
}
run default IPXGPU_X=1
run nofused IPXGPU_FUSED=0
run generic IPXGPU_SWEEP=generic
run plain IPXGPU_BAND_DEAL=plain
