"""stdin: one bench.py JSON line -> a short human summary (used by tools/gpu.sh)."""
import json
import sys

txt = sys.stdin.read().strip()
try:
    d = json.loads(txt.splitlines()[-1])
except Exception:
    print(txt[-2000:])
    sys.exit(0)
r = d.get("roofline") or {}
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "n_gpus", "dtype", "gpu_launches")})
print("e2e", d.get("e2e", {}).get("value"), "serial", d.get("e2e", {}).get("serial_value"), "| clocks", d.get("clocks"))
print("roofline", {k: r.get(k) for k in ("achieved", "achieved_executed", "frac", "frac_executed", "per_layer_roofline_frac", "kernel_share_of_step")})
print("latency", d.get("latency"), "| cpu", d.get("cpu_baseline"))
for row in (r.get("per_class") or [])[:40]:
    print("   ", row)
