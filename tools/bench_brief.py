"""Print the headline fields of a bench.py JSON line (and the per-kernel eager timings)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        with open(path) as f:
            d = json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    print("%s: value %.1f %s, e2e %.1f, ms/step %.2f, iter(graph) %.4f ms, clocks %s" % (
        path, d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], r.get("iteration_ms_graph", 0), d.get("clocks")))
    rows = [(r.get("kernel"), r.get("launch_ms"), r.get("launches_per_iteration"), r.get("frac"), r.get("share_of_iteration"))]
    rows += [(v["kernel"], v["launch_ms"], v["launches_per_iteration"], v["frac"], v["share_of_iteration"]) for v in (r.get("others") or {}).values()]
    for k, ms, n, frac, share in rows:
        if k:
            print("   %-34s %7.2f us x%d  hbm frac %.3f  share %.3f" % (k, ms * 1e3, n, frac, share))
    for key in ("setup_ms", "per_rank_ms_per_step", "reference_gpu", "issue_slots"):
        if key in d:
            print("  ", key, json.dumps(d[key])[:600])
