"""Prints the figures of a bench.py JSON line that matter when reading a GPU session's log.   python scripts/bench_brief.py FILE"""
import json
import sys

try:
    lines = [l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
except Exception as e:
    print("no bench line", e)
    sys.exit(0)
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "sharded_matches")}, d.get("stage_ms"))
e = d.get("e2e") or {}
print("   e2e", e.get("value"), e.get("ms_per_step"), e.get("one_panorama_at_a_time"), e.get("panoramas_in_flight") or e.get("two_panoramas_in_flight"), e.get("host_link_pinned_copy"))
r = d.get("roofline") or {}
print("   roofline", r.get("frac"), r.get("kernel_ms_per_step"), r.get("whole_step"))
for k in (r.get("kernels") or [])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("      ", k)
print("   clocks", d.get("clocks"), "cfg", (d.get("config") or {}).get("workload"))
