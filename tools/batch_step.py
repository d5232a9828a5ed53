#!/usr/bin/env python
"""A few steps of the batched engine (device-resident inputs) for ncu launch lists / captures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_b200 import api  # noqa: E402
import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "vlp16_100k"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda", 0)
    nseq = min(B, 8)
    setup = api.Context(0)
    seqs = []
    for s in range(nseq):
        mc, ms, scans = bench.make_inputs(workload, s, 2)
        setup.map_set_raw(mc, ms)
        mc_ds, ms_ds = setup.map_get_ds(0), setup.map_get_ds(1)
        seqs.append(dict(d_mc=torch.from_numpy(mc_ds).to(dev), d_ms=torch.from_numpy(ms_ds).to(dev),
                         d_scans=[(torch.from_numpy(sc.corner_last).to(dev), torch.from_numpy(sc.surf_last).to(dev),
                                   torch.from_numpy(sc.outlier_last).to(dev), init) for sc, init in scans]))
    setup.close()
    max_map = max(max(q["d_mc"].shape[0], q["d_ms"].shape[0]) for q in seqs) + 1000
    b = api.Batch(0, B, 8192, max_map)
    torch.cuda.synchronize()
    for i in range(steps):
        T = np.zeros((B, 6), np.float32)
        for s in range(B):
            q = seqs[s % nseq]
            c, s_, o, init = q["d_scans"][i % 2]
            b.scan_set_dev(s, c.data_ptr(), c.shape[0], s_.data_ptr(), s_.shape[0], o.data_ptr(), o.shape[0])
            b.map_set_ds_dev(s, q["d_mc"].data_ptr(), q["d_mc"].shape[0], q["d_ms"].data_ptr(), q["d_ms"].shape[0])
            T[s] = init
        Tres, st = b.register(T)
    print("iters", [x.iterations for x in st][:8], "device_ms", st[0].device_ms)
    b.close()


if __name__ == "__main__":
    main()
