"""Regenerates profiles/inst_counts.json: warp instructions (smsp__inst_executed.sum), DRAM bytes and ncu durations per
kernel for ONE step of every bench config, collected under ncu on a B200:

    python profiles/inst_counts.py            # runs ncu once per config (bench.py --one-step KEY), then writes the JSON
    python profiles/inst_counts.py --parse    # only re-parse gpurun_out/ic_<KEY>.csv

bench.py divides these counts by its own CUDA-event times to get the SM-issue fraction and prints "stale": true when
the hash of the CUDA sources stored here differs from the tree it runs on."""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = ["c0", "c1_bc3", "c1_bc5", "c2", "c3_sample", "c4"]
KNOWN = ["k_alpha_blocks", "k_alpha_optimal", "k_alpha_dxt3", "k_bc3_color", "k_bc1a_color", "k_bc1_icbc", "k_dxt1_quick", "k_bc6_rough",
         "k_bc6_tiles", "k_bc6_setup", "k_bc6_order", "k_bc6_search", "k_bc6_finish", "k_bc6_select", "k_bc7_rough", "k_bc7_tiles",
         "k_bc7_setup", "k_bc7_order", "k_bc7_search", "k_bc7_finish", "k_bc7_select", "k_set_image", "k_gamma", "k_box_down",
         "k_polyphase_x", "k_polyphase_y", "k_polyphase_2d", "k_normalize", "k_scale_bias", "k_grey_scale", "k_to_normal_map",
         "k_decode_blocks", "k_error_metric", "k_binarize", "k_quantize", "k_pixel_format", "k_xchg", "k_mip_encode"]
ALIAS = {"k_decode_dxt": "k_decode_blocks", "k_rgbm_alpha": "k_alpha_optimal", "k_dxt1g_optimal": "k_alpha_optimal", "k_export_rows": "k_xchg"}
METRICS = "smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"


def short_name(fn):
    i = fn.find("k_")
    if i < 0:
        return fn
    fn = fn[i:]
    for a, b in ALIAS.items():
        if fn.startswith(a):
            return b
    for k in sorted(KNOWN, key=len, reverse=True):
        if fn.startswith(k):
            return k
    return fn.split("<")[0].split("(")[0]


def parse(path):
    rows = {}
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        k = rows.setdefault(r["ID"], {"name": short_name(r["Kernel Name"])})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k[m] = v
    out = {}
    for k in rows.values():
        o = out.setdefault(k["name"], {"launches": 0, "warp_insts": 0.0, "dram_bytes": 0.0, "ncu_ms": 0.0})
        o["launches"] += 1
        o["warp_insts"] += k.get("smsp__inst_executed.sum", 0.0)
        o["dram_bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        o["ncu_ms"] += k.get("gpu__time_duration.sum", 0.0)
    return out


def main():
    outdir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    keys = [a for a in sys.argv[1:] if not a.startswith("--")] or KEYS
    if "--parse" not in sys.argv:
        for key in keys:
            log = os.path.join(outdir, "ic_%s.csv" % key)
            cmd = ["ncu", "--profile-from-start", "off", "--metrics", METRICS, "--clock-control", "none", "--csv", "--log-file", log,
                   sys.executable, os.path.join(ROOT, "bench.py"), "--one-step", key]
            print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True, cwd=ROOT)
    path = os.path.join(ROOT, "profiles", "inst_counts.json")
    try:
        doc = json.load(open(path))
    except Exception:
        doc = {"configs": {}}
    for key in keys:
        log = os.path.join(outdir, "ic_%s.csv" % key)
        if os.path.exists(log):
            doc["configs"][key] = {"kernels": parse(log)}
    doc["kernels_hash"] = bench.kernels_hash()
    doc["when"] = datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ")
    doc["how"] = "ncu --profile-from-start off --metrics %s --clock-control none, one step of bench.py --one-step KEY; per kernel: sum over its launches in the step" % METRICS
    json.dump(doc, open(path, "w"), indent=1, sort_keys=True)
    for key in keys:
        ks = doc["configs"].get(key, {}).get("kernels", {})
        tot = sum(v["ncu_ms"] for v in ks.values()) or 1.0
        print(key, {k: "%.3f ms %.1f%%" % (v["ncu_ms"], 100 * v["ncu_ms"] / tot) for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ncu_ms"])[:4]})


if __name__ == "__main__":
    main()
