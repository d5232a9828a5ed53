#!/usr/bin/env python
"""Per-stage device timings of the hot path (CUDA events on the context's stream, warm cache)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_b200 import api  # noqa: E402
import bench  # noqa: E402


def timed(stream, fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    stream.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    return e0.elapsed_time(e1) / reps, wall


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "vlp16_100k"
    print("host cores:", len(os.sched_getaffinity(0)), "workload:", workload)
    mc, ms, scans = bench.make_inputs(workload, 0, 2)
    ctx = api.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    dev = torch.device("cuda", 0)
    ctx.map_set_raw(mc, ms)
    mc_ds = ctx.map_get_ds(0); ms_ds = ctx.map_get_ds(1)
    d_mc = torch.from_numpy(mc_ds).to(dev); d_ms = torch.from_numpy(ms_ds).to(dev)
    d_mcr = torch.from_numpy(mc).to(dev); d_msr = torch.from_numpy(ms).to(dev)
    sc, init = scans[0]
    d_c = torch.from_numpy(sc.corner_last).to(dev); d_s = torch.from_numpy(sc.surf_last).to(dev)
    d_o = torch.from_numpy(sc.outlier_last).to(dev)
    d_T0 = torch.from_numpy(init.copy()).to(dev); d_T = torch.zeros(6, device=dev)
    print("map raw", mc.shape[0], ms.shape[0], "map ds", mc_ds.shape[0], ms_ds.shape[0],
          "scan", sc.corner_last.shape[0], sc.surf_last.shape[0], sc.outlier_last.shape[0])
    with torch.cuda.stream(stream):
        ctx.scan_set_dev(d_c.data_ptr(), d_c.shape[0], d_s.data_ptr(), d_s.shape[0], d_o.data_ptr(), d_o.shape[0])
        rows = []
        rows.append(("scan_set_dev (borrowed pointers)", timed(stream, lambda: ctx.scan_set_dev(
            d_c.data_ptr(), d_c.shape[0], d_s.data_ptr(), d_s.shape[0], d_o.data_ptr(), d_o.shape[0]))))
        rows.append(("downsampleCurrentScan (4 voxel filters)", timed(stream, lambda: ctx.downsample_current_scan(False))))
        rows.append(("map_set_ds_dev (2 index builds)", timed(stream, lambda: ctx.map_set_ds_dev(
            d_mc.data_ptr(), d_mc.shape[0], d_ms.data_ptr(), d_ms.shape[0]))))

        def s2m():
            d_T.copy_(d_T0)
            ctx.s2m_optimize_dev(d_T.data_ptr())
        rows.append(("scan2MapOptimization loop (prepare + persistent kernel)", timed(stream, s2m)))
        rows.append(("map_set_raw_dev (2 map voxel filters + 2 index builds)", timed(stream, lambda: ctx.map_set_raw_dev(
            d_mcr.data_ptr(), d_mcr.shape[0], d_msr.data_ptr(), d_msr.shape[0]), reps=10, warm=2)))
        ctx.map_set_ds_dev(d_mc.data_ptr(), d_mc.shape[0], d_ms.data_ptr(), d_ms.shape[0])
        ms_it, nq = ctx.s2m_time_iteration(init, reps=50)
        T, st = ctx.s2m_optimize(init)
    for name, (dms, wall) in rows:
        print(f"{name:62s} device {dms * 1e3:9.1f} us   host-wall {wall * 1e3:9.1f} us")
    print(f"one accumulate-only iteration: {ms_it * 1e3:.1f} us for {nq} queries; full optimize stats: {st.as_dict()}")
    cp = ctx.s2m_get_cta_profile()
    if cp.shape[0]:
        nc_ctas = int(np.ceil(st.n_corner_ds / max(1, int(np.ceil((st.n_corner_ds + st.n_surf_ds) / cp.shape[0])))))
        for name, sel in (("corner CTAs", slice(0, max(nc_ctas - 1, 1))), ("surf CTAs", slice(nc_ctas, None))):
            q = cp[sel]
            if q.shape[0]:
                print(f"per-CTA cycles, last iteration, {name} ({q.shape[0]}): "
                      + ", ".join(f"{n} med {np.median(q[:, k]):.0f} max {q[:, k].max():.0f}" for k, n in enumerate("ABCW")))
    for it in range(st.iterations):
        print(f"CTA-0 cycles per phase, iteration {it}:", ctx.s2m_get_profile(it))
    # odometry
    from lego_loam_b200 import synth
    w = synth.make_world()
    od = synth.make_odometry_pair(w, synth.VLP16, np.array([0, 1.0, 0, 3.0, 0, -4.0]),
                                  np.array([0.002, 0.015, -0.001, 0.01, 0.005, -0.15]), seed=1)
    ctx.odom_set_last(od.corner_last, od.surf_last); ctx.odom_set_features(od.corner_sharp, od.surf_flat)
    for _ in range(3):
        T, s0, s1 = ctx.odom_optimize(np.zeros(6, np.float32))
    print(f"updateTransformation: device {s0.device_ms * 1e3:.1f} us, iters surf {s0.iterations} corner {s1.iterations}, "
          f"sharp {od.corner_sharp.shape[0]} flat {od.surf_flat.shape[0]} last {od.corner_last.shape[0]}/{od.surf_last.shape[0]}")
    ctx.close()


if __name__ == "__main__":
    main()
