"""Print the few numbers of a bench.py JSON line that matter while iterating."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print("value %.2f %s  ms/step %.1f  e2e %.2f  launches %s  roofline %.3f (%.0f TF/s, share %.2f)  clocks %s" % (
        d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value", float("nan")), d.get("gpu_launches"),
        r.get("frac") or 0, r.get("achieved") or 0, r.get("share_of_step") or 0, (d.get("clocks") or {}).get("sm_mhz")))
    for k in ("strict_bf16x3", "fast_tail", "pano768"):
        if k in d:
            print("  %s: %s" % (k, {a: b for a, b in d[k].items() if a in ("value", "ms_per_step", "gemm_algorithmic_tflops")}))
    if "train" in d and d["train"]:
        t = d["train"]
        print("  train: %.1f %s ms %.1f parts %s" % (t["value"], t["unit"], t["ms_per_step"], t["config"]["part_ms"]))
    if "call_ms" in d:
        print("  call_ms:", d["call_ms"])
    for s in (r.get("by_shape") or [])[:26]:
        print("   %-60s %8.2f ms %4d  %7.1f TF/s" % (s["shape"], s["ms_per_step"], s["launches_per_step"], s["algorithmic_tflops"] or 0))
