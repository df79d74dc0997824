"""Device-resident and end-to-end (pinned host in, host out) time of bench.py's per-config records (developer tool):
    python profiles/time_e2e_configs.py [c0 c1_bc3 c1_bc5 c3_sample c4]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nvtt_b200_loader  # noqa: E402

m = nvtt_b200_loader.load()
ctx = m.Context(0)
keys = sys.argv[1:]
for spec in bench.config_specs(m):
    if keys and spec["key"] not in keys:
        continue
    r = bench.run_config(m, ctx, spec, {}, True, 1965.0)
    print("%-10s device %9.3f ms   e2e %9.3f ms" % (spec["key"], r["ms_per_step"], r["e2e"]["ms_per_step"]), flush=True)
