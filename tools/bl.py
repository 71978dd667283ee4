"""Print the key numbers of a bench.py JSON line (last line of a log file)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(path, "ms/step %.4f value %.0f | roi %.4f ms frac %.3f | serial %s | host %.3f | e2e %.0f (%.2f ms) | pipelined %s | %s" % (
        d["ms_per_step"], d["value"], r.get("kernel_ms_mean", 0), r.get("frac", 0),
        {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, d.get("host_enqueue_ms_per_step", 0),
        d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), (d.get("pipelined") or {}).get("ms_per_step"), d["config"].get("roi_align_mode")))
