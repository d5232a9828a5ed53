#!/usr/bin/env python
"""Feature extraction of a few sweeps (for ncu launch lists / timing): python tools/features_step.py [sensor] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_b200 import api, synth  # noqa: E402

sensor = synth.SENSORS[sys.argv[1] if len(sys.argv) > 1 else "vlp16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = synth.make_world()
sws = [synth.make_segmented_sweep(w, sensor, [0, 0.05 + 0.01 * k, 0, 3 + 0.4 * k, 0, 5], 11 + k) for k in range(2)]
c = api.Context(0); c.features_init(sensor.n_scan, sensor.horizon)
ms = []
for i in range(reps):
    t0 = time.perf_counter(); counts, d = c.features_extract(sws[i % 2]); ms.append(((time.perf_counter() - t0) * 1e3, d))
print(sensor.name, "points", [s.cloud.shape[0] for s in sws], "counts", counts)
print("wall ms", np.round([m[0] for m in ms], 3).tolist())
print("slowest ring", c.features_get_profile())
print("device ms", np.round([m[1] for m in ms], 3).tolist())
c.close()
