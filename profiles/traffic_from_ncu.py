#!/usr/bin/env python
"""Summarise an ncu launch list into profiles/r2_traffic.json (read by bench.py for `roofline.traffic`).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
        --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
        python profiles/prof_forward.py --chunks 1 --reps 1 --precision fp16x2          # on the GPU box
    python profiles/traffic_from_ncu.py gpurun_out/r2_launches.csv profiles/r2_traffic.json 37888   # here

Per-launch times under ncu are cold-cache and serialised: the SHARES are what is comparable with the timed step.
The first occurrence of the forward in the list is a warm-up pass (prof_forward.py runs the chunk three times +
one profiled pass), so only the LAST `launches_per_pass` launches of each kernel are kept."""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    src, dst, chunk = sys.argv[1], sys.argv[2], int(sys.argv[3])
    passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # how many forward passes the capture holds
    rows = []
    with open(src, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("dsb::", "").replace("(anonymous namespace)::", "")
        k = per.setdefault((r["ID"], name), {})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            k["time_us"] = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            k[m] = val * scale
    kern = OrderedDict()
    for (_, name), v in per.items():
        if "at::" in name:              # torch's own kernels (input generation of the driver script), not the library's
            continue
        kern.setdefault(name.replace("void ", "").replace("<unnamed>::", ""), []).append(v)
    out = {"source": src + " (ncu --clock-control none)", "chunk_samples": chunk, "kernels": OrderedDict()}
    total = 0.0
    for name, launches in kern.items():
        keep = launches[-(len(launches) // passes):] if passes > 1 else launches
        t = sum(x.get("time_us", 0.0) for x in keep)
        dram = sum(x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0) for x in keep)
        l2 = sum(x.get("lts__t_bytes.sum", 0.0) for x in keep)
        out["kernels"][name] = {"launches": len(keep), "time_us": t, "dram_bytes": dram, "l2_bytes": l2,
                                "dram_TBps": dram / t / 1e6 if t else 0.0}
        total += t
    for v in out["kernels"].values():
        v["share"] = v["time_us"] / total if total else 0.0
    out["total_time_us"] = total
    out["dram_bytes_per_sample"] = sum(v["dram_bytes"] for v in out["kernels"].values()) / chunk
    out["algorithmic_bytes_per_sample"] = 1724
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: (round(v["time_us"], 1), round(v["dram_bytes"] / 1e6, 1)) for k, v in out["kernels"].items()}))


if __name__ == "__main__":
    main()
