#!/usr/bin/env python
"""Feature extraction, batched: S sequences, one sweep each per step (python tools/features_batch_bench.py [S] [steps])."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_b200 import api, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = synth.make_world()
base = [synth.make_segmented_sweep(w, synth.VLP16, [0, 0.05 + 0.01 * k, 0, 3 + 0.4 * k, 0, 5], 11 + k) for k in range(4)]


class Held:          # sweep with its 32 B-stride cloud built once (the caller's PCL cloud)
    def __init__(self, sw):
        self.__dict__.update(sw.__dict__); self.cloud32 = api.to_pcl(sw.cloud)


held = [Held(s) for s in base]
b = api.Batch(0, S, 4096, 4096); b.features_init(16, 1800)
packs = [b.features_pack([held[(s + i) % 4] for s in range(S)]) for i in range(2)]
dev, wall = [], []
for i in range(steps + 3):
    t0 = time.perf_counter(); counts, ms = b.features_extract(packs[i % 2]); t1 = time.perf_counter()
    if i >= 3:
        dev.append(ms); wall.append((t1 - t0) * 1e3)
print(f"S={S}: device {np.median(dev):.3f} ms/step = {S / np.median(dev) * 1e3:.0f} sweeps/s; "
      f"C ABI wall {np.median(wall):.3f} ms/step = {S / np.median(wall) * 1e3:.0f} sweeps/s; counts[0]={counts[0].tolist()}")
b.close()
