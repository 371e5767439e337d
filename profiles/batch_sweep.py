#!/usr/bin/env python
"""SURVEY section 8(d) cfg 5: the forward over batch sizes 512 ... 262,144 (device-resident inputs, one dsb_forward_dev
call per batch).

  python profiles/batch_sweep.py [precision]            timing sweep (best of a few repetitions) -> one JSON line per batch
  python profiles/batch_sweep.py [precision] --once      exactly ONE forward per batch size after a warm-up, for ncu:

      ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
          --clock-control none --csv --log-file gpurun_out/r2_sweep_ncu.csv python profiles/batch_sweep.py fp16x2 --once
      python profiles/batch_sweep.py --merge gpurun_out/r2_sweep_ncu.csv gpurun_out/r2_sweep.jsonl profiles/r2_batch_sweep_fp16x2.jsonl

  --merge adds, per batch size, achieved HBM GB/s (ncu dram bytes / summed kernel time) and the time-weighted tensor-pipe
  utilisation of the launches of that batch's forward (ncu counters; per-launch times are cold-cache and serialised)."""
import csv
import json
import os
import re
import sys

SIZES = [512 << i for i in range(10)]       # 512 ... 262,144
WARM = 4096


def merge(ncu_csv, sweep_jsonl, out_path):
    with open(ncu_csv, newline="") as fh:
        rows = list(csv.DictReader([ln for ln in fh if not ln.startswith("==")]))
    launches, order = {}, []
    for r in rows:
        lid = int(r["ID"])
        if lid not in launches:
            launches[lid] = {"name": re.sub(r"\(.*", "", r["Kernel Name"])}
            order.append(lid)
        v = float(r["Metric Value"].replace(",", ""))
        u, m = r["Metric Unit"], r["Metric Name"]
        if m == "gpu__time_duration.sum":
            launches[lid]["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif m.startswith("dram__bytes"):
            launches[lid]["dram"] = launches[lid].get("dram", 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        elif m.startswith("sm__pipe_tensor"):
            launches[lid]["tensor_pct"] = v
    seq = [launches[i] for i in order]
    # a forward = everything from one lstm_feat_kernel launch up to (not including) the next one
    starts = [i for i, l in enumerate(seq) if "lstm_feat_kernel" in l["name"]]
    groups = [seq[a:b] for a, b in zip(starts, starts[1:] + [len(seq)])]
    timing = {}
    if sweep_jsonl and os.path.exists(sweep_jsonl):
        for ln in open(sweep_jsonl):
            if ln.startswith("{"):
                d = json.loads(ln)
                timing[d["batch"]] = d
    # --once runs: warm-up forward(s) first, then one forward per size; chunked sizes span several feat launches
    sizes = list(SIZES)
    chunks = [max(1, -(-n // 37888)) for n in sizes]
    need = sum(chunks)
    groups = groups[len(groups) - need:]
    with open(out_path, "w") as out:
        gi = 0
        for n, c in zip(sizes, chunks):
            g = [l for grp in groups[gi:gi + c] for l in grp]
            gi += c
            us = sum(l.get("us", 0.0) for l in g)
            dram = sum(l.get("dram", 0.0) for l in g)
            tens = sum(l.get("tensor_pct", 0.0) * l.get("us", 0.0) for l in g) / us if us else 0.0
            rec = dict(timing.get(n, {"batch": n}))
            rec.update({"ncu_kernel_time_us": round(us, 1), "ncu_dram_bytes": dram, "ncu_hbm_GBps": round(dram / us / 1e3, 1) if us else 0.0,
                        "ncu_tensor_pipe_pct_time_weighted": round(tens, 2), "ncu_launches": len(g),
                        "dram_bytes_per_sample": round(dram / n, 1)})
            out.write(json.dumps(rec) + "\n")
            print(json.dumps(rec))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--merge":
        return merge(sys.argv[2], sys.argv[3], sys.argv[4])
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from deepsignal_b200.engine import Engine
    from deepsignal_b200.weights import random_checkpoint
    prec = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "fp16x2"
    once = "--once" in sys.argv
    eng = Engine(0)
    eng.load_weights(random_checkpoint(17))
    nmax = SIZES[-1]
    g = torch.Generator(device="cuda").manual_seed(1)
    kmer = torch.randint(0, 4, (nmax, 17), device="cuda", dtype=torch.int32, generator=g)
    means = torch.randn(nmax, 17, device="cuda", generator=g)
    stds = torch.rand(nmax, 17, device="cuda", generator=g)
    lens = torch.randint(1, 60, (nmax, 17), device="cuda", generator=g).float()
    signals = torch.randn(nmax, 360, device="cuda", generator=g)
    if once:
        eng.set_profiling(True)      # single stream: the launch list of a forward is contiguous
        for n in [WARM] + SIZES:
            eng.forward_torch(kmer[:n], means[:n], stds[:n], lens[:n], signals[:n], precision=prec)
            torch.cuda.synchronize()
        return
    for n in SIZES:
        run = lambda: eng.forward_torch(kmer[:n], means[:n], stds[:n], lens[:n], signals[:n], precision=prec)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        best = 1e30
        reps = max(3, min(50, (1 << 19) // n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        print(json.dumps({"batch": n, "precision": prec, "ms": best, "samples_per_s": n / best * 1e3,
                          "tensor_frac_of_sustained_peak_algorithmic": n / best * 1e3 * 198.61e6 / 1403e12}))


if __name__ == "__main__":
    main()
