# the two arms of the driver's round-end bench, N = 1, default flags
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
tail -c 400 gpurun_out/bench_default.err
( time python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.json", "gpurun_out/bench_reference.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, {k: d.get(k) for k in ("value", "unit", "ms_per_step", "gpu_launches", "vs_baseline")})
            print("  e2e", d.get("e2e")); print("  roofline", {k: v for k, v in (d.get("roofline") or {}).items() if k != "ncu"})
            print("  cpu_baseline", d.get("cpu_baseline")); print("  clocks", d.get("clocks"))
            o = d.get("others") or {}
            for k, v in o.items():
                if isinstance(v, dict): print("  others", k, {kk: vv for kk, vv in v.items() if isinstance(vv, (int, float, str)) and kk != "workload"})
PY
