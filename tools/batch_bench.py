#!/usr/bin/env python
"""Throughput of the batched multi-registration engine (llb_batch_*) for several batch sizes, with the per-stage
CUDA-event split of one step; device-resident inputs and host-cloud inputs."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_b200 import api  # noqa: E402
import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "vlp16_100k"
    sizes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "16", "64"])]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    dev = torch.device("cuda", 0)
    nseq = min(max(sizes), 8)                 # distinct sequences; slots beyond reuse them
    setup = api.Context(0)
    seqs = []
    for s in range(nseq):
        mc, ms, scans = bench.make_inputs(workload, s, 2)
        setup.map_set_raw(mc, ms)
        seqs.append(dict(mc_ds=setup.map_get_ds(0), ms_ds=setup.map_get_ds(1), scans=scans))
    setup.close()
    for q in seqs:
        q["mc32"] = api.to_pcl(q["mc_ds"]); q["ms32"] = api.to_pcl(q["ms_ds"])
        q["scans32"] = [(api.to_pcl(sc.corner_last), api.to_pcl(sc.surf_last), api.to_pcl(sc.outlier_last), init) for sc, init in q["scans"]]
        q["d_mc"] = torch.from_numpy(q["mc_ds"]).to(dev); q["d_ms"] = torch.from_numpy(q["ms_ds"]).to(dev)
        q["d_scans"] = [(torch.from_numpy(sc.corner_last).to(dev), torch.from_numpy(sc.surf_last).to(dev),
                         torch.from_numpy(sc.outlier_last).to(dev), init) for sc, init in q["scans"]]
    max_map = max(max(q["mc_ds"].shape[0], q["ms_ds"].shape[0]) for q in seqs) + 1000
    for B in sizes:
        prm = api.default_params(); prm.pin_host_clouds = 1
        b = api.Batch(0, B, 8192, max_map, prm)
        stream = torch.cuda.ExternalStream(b.stream, device=0)

        def step_dev(i, set_map=True):
            T = np.zeros((B, 6), np.float32)
            for s in range(B):
                q = seqs[s % nseq]
                c, s_, o, init = q["d_scans"][i % 2]
                b.scan_set_dev(s, c.data_ptr(), c.shape[0], s_.data_ptr(), s_.shape[0], o.data_ptr(), o.shape[0])
                if set_map:
                    b.map_set_ds_dev(s, q["d_mc"].data_ptr(), q["d_mc"].shape[0], q["d_ms"].data_ptr(), q["d_ms"].shape[0])
                T[s] = init
            b.register_async(T)

        def step_host(i, set_map=True):
            T = np.zeros((B, 6), np.float32)
            for s in range(B):
                q = seqs[s % nseq]
                c, s_, o, init = q["scans32"][i % 2]
                b.scan_set_pcl(s, c, s_, o)
                if set_map:
                    b.map_set_ds_pcl(s, q["mc32"], q["ms32"])
                T[s] = init
            b.register_async(T)

        out = {"B": B, "workload": workload}
        for name, fn, sm in (("dev", step_dev, True), ("dev_resident_map", step_dev, False), ("host", step_host, True),
                             ("host_resident_map", step_host, False)):
            for i in range(3):
                fn(i, True if i == 0 else sm); Tres, st = b.result()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dms = 0.0
            for i in range(steps):
                fn(i, sm); Tres, st = b.result()
                dms += st[0].device_ms
            wall = time.perf_counter() - t0
            out[name] = {"reg_per_s_wall": B * steps / wall, "ms_per_step_wall": wall / steps * 1e3,
                         "ms_per_step_device": dms / steps, "reg_per_s_device": B * steps / (dms * 1e-3)}
        b.set_profile(True)
        step_dev(0, True); Tres, st = b.result()
        prof, geo = b.get_profile()
        out["profile_ms"] = prof; out["geometry"] = geo
        out["iters"] = [x.iterations for x in st][:8]
        out["launches_per_step"] = None
        l0 = b.launch_count(); step_dev(1, True); b.result(); out["launches_per_step"] = b.launch_count() - l0
        print(json.dumps(out))
        b.close()


if __name__ == "__main__":
    main()
