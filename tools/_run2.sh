IPXGPU_FUSED_TRACE=1 timeout 300 python tools/profile_apply.py --reps 5 --pcr 6 2>&1 | grep -E "fused trace|errflag"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused6.json 2> gpurun_out/bench_fused6.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_fused6.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "apply_us", d["roofline"]["apply_us_in_loop"])
PY
tail -5 gpurun_out/bench_fused6.err
