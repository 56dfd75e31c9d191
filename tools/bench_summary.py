"""Print the headline fields of bench.py JSON lines: python tools/bench_summary.py file..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        line = [x for x in open(f) if x.startswith("{")][-1]
        d = json.loads(line)
        lb = d["roofline"]["latency_bound"]
        e = d.get("e2e") or {}
        print(f, "value %.3e" % d["value"], "ms/step %.1f" % d["ms_per_step"], "units", lb.get("units"), "seg", lb.get("segment_sites"),
              "ctas", lb.get("resident_ctas"), "us/site/cta %.2f" % lb.get("us_per_site_per_cta", 0), "e2e %.3e" % e.get("value", 0),
              lb.get("resampling_sites"))
    except Exception as ex:  # noqa: BLE001
        print(f, "unreadable:", ex)
