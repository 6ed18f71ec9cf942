"""One line per case of a tools/gpu_probe.py log: status and timings."""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        keys = ("ms", "tflops", "fwd_ms", "bwd_ms", "fwd_tflops_causal", "bwd_tflops_causal", "gbs", "fwd_gbs", "bwd_gbs")
        t = {k: round(v, 4) for k, v in d.items() if k in keys and isinstance(v, float)}
        if t or not d.get("ok"):
            print(d["case"], "ok" if d.get("ok") else "FAIL " + str(d)[:300], t)
