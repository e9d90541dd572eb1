"""One full wave (148 blocks x 256 query rows) of the wide-row first tier against a 500k x 2000 database: the launch
the round-2 ncu capture of the K-loop kernel profiles (C5-shaped data, generated on the device)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_cells_device  # noqa: E402
from vivae_b200.device import knn_device  # noqa: E402

dev = torch.device("cuda", 0)
x = synth_cells_device(500_000, 2000, dev)
for _ in range(int(os.environ.get("REPS", "2"))):
    st = {}
    knn_device(x, 50, 0, int(os.environ.get("NQ_BLOCKS", "148")) * 256, stats=st)
    print(json.dumps(st))
