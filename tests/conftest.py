import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    arrays = dict(np.load(os.path.join(GOLDEN, "features_ref.npz")))
    with open(os.path.join(GOLDEN, "features_ref.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def golden_reads(golden):
    from deepsignal_b200.synth import ReadBatch
    a, m = golden
    return ReadBatch(a["reads_raw"], a["reads_raw_off"], a["reads_scaling"], a["reads_offset"],
                     a["reads_ev_start"], a["reads_ev_len"], a["reads_ev_base"], a["reads_ev_off"],
                     a["reads_chrom_start"], a["reads_strand_minus"], a["reads_chrom_len"],
                     m["reads_chrom"], m["reads_id"], m["reads_strand"])
