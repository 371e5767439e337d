#!/usr/bin/env python
"""One configs[1] extraction (4,000 synthetic reads, device-resident) after a warm-up, for an ncu launch list:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/r2_launches_extract.csv python profiles/prof_extract.py

Prints the CUDA-event time of the un-profiled call and the library's per-kernel profile."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_b200.engine import Engine          # noqa: E402
from deepsignal_b200.synth import synth_reads, SEED_CFG2  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    eng = Engine(0)
    eng.kmer_len, eng.signal_len = 17, 360
    rb = synth_reads(n, seed=SEED_CFG2)
    dev = torch.device("cuda", 0)
    dr = eng.reads_to_device(rb, dev)
    bufs = eng.alloc_site_bufs(max(1024, int(rb.ev_off[-1] * 0.03)), dev)
    eng.extract_torch(dr, rb.n_reads, bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ns = eng.extract_torch(dr, rb.n_reads, bufs)
    e1.record()
    torch.cuda.synchronize()
    eng.set_profiling(True)
    eng.extract_torch(dr, rb.n_reads, bufs)
    torch.cuda.synchronize()
    prof = {k: round(v["ms"], 4) for k, v in eng.profile().items() if k.startswith("x_")}
    n_raw, n_ev = int(rb.raw_off[-1]), int(rb.ev_off[-1])
    print(json.dumps({"reads": n, "sites": ns, "raw_samples": n_raw, "bases": n_ev, "ms": e0.elapsed_time(e1),
                      "algorithmic_bytes": 2 * n_raw + 9 * n_ev + 1677 * ns, "per_kernel_ms": prof}))


if __name__ == "__main__":
    main()
