#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures of bench.py (one per config, tools/r02_evidence.sh):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the traversal / deposit kernels, averaged over
the captured launches of each kernel class, with the capture it came from and the commit it was taken at.

  python tools/make_traffic.py cfg2=gpurun_out/r02p_cfg2.ncu-rep cfg5=gpurun_out/r02p_cfg5.ncu-rep ...
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def klass(name):
    if "k_trace_wide<1" in name or "k_trace_wide<(bool)1" in name or "k_trace_fast<1" in name or "k_trace_hw1_fast<1" in name \
            or "k_trace_hw1_fast<(bool)1" in name or "k_trace_fast<(bool)1" in name:
        return "shadow"
    if "k_trace_wide<" in name or "k_trace_fast<" in name or "k_trace_hw1_fast<" in name:
        return "extend"
    if "k_ppm_photon_shade" in name:
        return "deposit"
    return None


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    commit = os.environ.get("B200RT_COMMIT") or subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT,
                                                             capture_output=True, text=True).stdout.strip()
    path0 = os.path.join(ROOT, "profiles", "traffic.json")
    res = json.load(open(path0)) if (os.environ.get("B200RT_TRAFFIC_MERGE") and os.path.exists(path0)) else {}
    res.update({"_source": f"ncu --set full --clock-control none captures of bench.py at commit {commit} (tools/r02_evidence.sh); "
                      "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of a class"})
    for arg in sys.argv[1:]:
        cfg, path = arg.split("=", 1)
        agg = {}
        for d, u in rows_of(path):
            k = klass(d["Kernel Name"])
            if not k:
                continue
            b = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            t = float(d["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u["gpu__time_duration.sum"], 1e-9)
            a = agg.setdefault(k, {"kernel": d["Kernel Name"].split("(")[0], "bytes": [], "secs": []})
            a["bytes"].append(b)
            a["secs"].append(t)
        res[cfg] = {k: {"kernel": a["kernel"], "dram_bytes_per_launch": sum(a["bytes"]) / len(a["bytes"]),
                        "launches_captured": len(a["bytes"]),
                        "mean_launch_ms_under_ncu": 1e3 * sum(a["secs"]) / len(a["secs"]),
                        "report": os.path.basename(path)} for k, a in agg.items()}
    json.dump(res, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
