#!/usr/bin/env python
"""bench.py -- scan-to-map registrations/s on B200 (BASELINE.json metric, config 5: 64 independent VLP-16 sequences).

A REGISTRATION is one mapping cycle of the reference node for one sequence (the north-star path end to end):
    local map of the sequence assembled from its key-frames     (transformPointCloud + concatenation, MO:1033-1056)
    two map voxel filters                                        (MO:1057-1064)
    spatial index of the two voxel-DS maps                       (replaces kdtree->setInputCloud x2, MO:1333-1334)
    downsampleCurrentScan                                        (MO:1067-1091)
    scan2MapOptimization                                         (MO:1329-1350: <= 10 LM iterations)
A STEP registers one new sweep for each of the S sequences of a GPU with the batched engine (llb_batch_*: every kernel
launch covers all slots, NB batches alternate so that the host work of one overlaps the kernels of the other); the
key-frame clouds of every sequence are device-resident (llb_batch_keyframe_add: each was uploaded once, when it was the
current sweep).  Ranks are replicas (weak scaling, no data-path collective).

  value : the new sweeps already in HBM; CUDA events (first start -> last end over the NB batch streams), max over ranks.
  e2e   : the same steps through the C ABI with HOST sweeps in pcl::PointXYZI layout: H2D of the sweep and D2H of pose +
          stats inside the timed region, wall clock, one host thread per batch, on EVERY rank.
  roofline : the registration kernel (kNN + fits + LM steps of all iterations, ONE launch per step), 96 algorithmic bytes
          per query-iteration, timed with CUDA events on the batch stream; roofline_stages: index build
          and map voxel filters against their own algorithmic bytes.
  cpu_baseline : the reference's own statements for the same cycle (oracle/_ref: the unmodified mapOptmization.cpp; kind
          "reference") on a bounded sample, 1 core.
  registration_only : secondary (rank 0): the drop-in signature alone - DS map handed over, downsampleCurrentScan +
          scan2MapOptimization - device-resident and with host clouds (scan AND DS map over PCIe), as round 1 reported.
  latency : one sequence alone on the single-registration path (one persistent kernel), L2 flushed.
  --impl reference : the same cycle with the reference's CPU implementation on all host cores (one stream per core).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_QUERY = 96          # 16 B query float4 + 5 x 16 B neighbour float4
WORKLOADS = {
    # name: (sensor, n_corner_raw, n_surf_raw, corner radius, surf radius)   -> DS map size
    "vlp16_100k": ("vlp16", 400000, 110000, 120.0, 60.0),
    "vlp16_50k": ("vlp16", 150000, 50000, 90.0, 45.0),
    "hdl32e_300k": ("hdl32e", 900000, 420000, 160.0, 110.0),
    "vls128_2m": ("vls128", 1500000, 6500000, 260.0, 200.0),
}


def make_inputs(workload: str, seq_id: int, n_scans: int):
    """One local map + n_scans sweeps around it (an independent sequence)."""
    from lego_loam_b200 import synth
    sensor, ncr, nsr, rad, srad = WORKLOADS[workload]
    w = synth.make_world(synth.SEED0 + seq_id)
    rng = np.random.default_rng(7000 + seq_id)
    centre = np.array([rng.uniform(-10, 10), 0.0, rng.uniform(-10, 10)])
    mc, ms = synth.make_local_map(w, centre, ncr, nsr, seed=11 + seq_id, radius=rad, surf_radius=srad)
    scans = []
    for k in range(n_scans):
        pose = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-3.1, 3.1), rng.uniform(-0.02, 0.02),
                         centre[0] + rng.uniform(-8, 8), rng.uniform(-0.03, 0.03), centre[2] + rng.uniform(-8, 8)])
        sc = synth.make_mapping_scan(w, synth.SENSORS[sensor], pose, seed=100 * seq_id + k)
        scans.append((sc, synth.perturb_pose(pose, rng)))
    return mc, ms, scans


def sample_clocks(stop_evt, out, dev):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"
    while not stop_evt.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(dev)],
                               capture_output=True, text=True, timeout=5).stdout.strip()
            if r:
                out.append([x.strip() for x in r.split(",")])
        except Exception:
            pass
        stop_evt.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit())
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for k, nme in enumerate(names):
            if len(s) > 3 + k and s[3 + k].lower().startswith("active"):
                reasons.add(nme)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def cpu_registration_factory(mc_ds, ms_ds, scans):
    """Returns (fn(i) -> pose, kind) running the reference CPU path on registration i."""
    kind = "port"
    ref_harness = None
    try:
        from oracle import ref_harness                       # reference-linked harness, when built
        if ref_harness.available():
            kind = "reference"
    except Exception:
        ref_harness = None
    if kind == "reference":
        mo = ref_harness.MapOptimization()
    else:
        import oracle
        oracle.set_trig_mode(0)
        mo = oracle.MapOptimization()

    def run(i):
        sc, init = scans[i % len(scans)]
        mo.set_map_ds(mc_ds, ms_ds)
        mo.set_scan(sc.corner_last, sc.surf_last, sc.outlier_last)
        mo.transformTobeMapped = init
        mo.downsampleCurrentScan()
        mo.scan2MapOptimization()
        return mo.transformTobeMapped
    return run, kind


_T0 = time.time()


def _log(msg):
    print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def make_cycle_sequence(args):
    """One synthetic sequence of the mapping cycle: n_kf key-frame sweeps ~1 m apart along a gently turning path (the
    reference thins key poses to 1 m, MO:1011-1012) and n_new further sweeps beside the path -> (key poses (n_kf, 6) f32,
    [(corner, surf, outlier)] key-frame sweeps, [(corner, surf, outlier, true pose)] new sweeps)."""
    seq_id, n_kf, n_new, sensor = args
    from lego_loam_b200 import synth
    w = synth.make_world(synth.SEED0 + 500 + seq_id)
    rng = np.random.default_rng(8000 + seq_id)
    yaw0 = rng.uniform(-3.0, 3.0)
    x0, z0 = rng.uniform(-25, 25, 2)
    poses, kf, new = [], [], []
    for k in range(n_kf + n_new):
        j = k if k < n_kf else n_kf // 2 + 3 * (k - n_kf)
        yaw = yaw0 + 0.01 * j
        pose = np.array([0.004 * np.sin(0.3 * j), yaw, 0.004 * np.cos(0.2 * j),
                         x0 + 1.0 * j * np.sin(yaw0 + 0.005 * j), 0.0, z0 + 1.0 * j * np.cos(yaw0 + 0.005 * j)])
        if k >= n_kf:
            pose[3] += 0.35; pose[5] += 0.2                  # the new sweeps are not on a key-frame
        sc = synth.make_mapping_scan(w, synth.SENSORS[sensor], pose, seed=9000 + 1000 * seq_id + k)
        if k < n_kf:
            poses.append(pose.astype(np.float32)); kf.append((sc.corner_last, sc.surf_last, sc.outlier_last))
        else:
            new.append((sc.corner_last, sc.surf_last, sc.outlier_last, pose))
    return np.stack(poses), kf, new


def make_cycle_sequences(ids, n_kf, n_new, sensor):
    """the sequences are ray-cast on the host (0.25 s per sweep): all cores, untimed set-up"""
    import multiprocessing as mp
    jobs = [(i, n_kf, n_new, sensor) for i in ids]
    cores = max(1, min(len(jobs), len(os.sched_getaffinity(0))))
    if cores == 1:
        return [make_cycle_sequence(j) for j in jobs]
    with mp.get_context("spawn").Pool(cores) as pool:
        return pool.map(make_cycle_sequence, jobs)


def cpu_raw_map(key_poses, kf):
    """Raw local map of a sequence on the host (reference arm / cpu_baseline without a device): every key-frame's DS
    clouds (downsampleCurrentScan of the sweep, what saveKeyFramesAndFactor stores, MO:1443-1453) moved by its key pose
    (transformPointCloud MO:545-575) and concatenated (MO:1050-1054)."""
    import oracle
    from lego_loam_b200 import synth
    oracle.set_trig_mode(0)
    mo = oracle.MapOptimization()
    rc, rs = [], []
    for pose, (c, s, o) in zip(key_poses, kf):
        mo.set_scan(c, s, o); mo.downsampleCurrentScan()
        for cloud, dst in ((mo.scan_ds(0), rc), (mo.scan_ds(1), rs), (mo.scan_ds(2), rs)):
            out = cloud.copy()
            out[:, :3] = synth.apply_pose(pose.astype(np.float64), cloud[:, :3].astype(np.float64)).astype(np.float32)
            dst.append(out)
    return np.ascontiguousarray(np.concatenate(rc)), np.ascontiguousarray(np.concatenate(rs))


def cpu_cycle_factory(raw_c, raw_s, new, inits):
    """-> (fn(i) -> pose, kind): the reference's statements for one mapping cycle on the CPU (raw local map given)"""
    kind, mo = "port", None
    try:
        from oracle import ref_harness
        if ref_harness.available():
            kind = "reference"; mo = ref_harness.MapOptimization()
    except Exception:
        mo = None
    if mo is None:
        import oracle
        oracle.set_trig_mode(0); mo = oracle.MapOptimization()

    def run(i):
        c, s_, o, _ = new[i % len(new)]
        mo.set_map_raw(raw_c, raw_s)                         # MO:1057-1064: the two map voxel filters
        mo.set_scan(c, s_, o)
        mo.transformTobeMapped = inits[i % len(new)]
        mo.downsampleCurrentScan()
        mo.scan2MapOptimization()
        return mo.transformTobeMapped
    return run, kind


_REF_BARRIER = None


def _ref_init(barrier):
    global _REF_BARRIER
    _REF_BARRIER = barrier


def _ref_worker(args):
    seq_id, n_kf, n_new, sensor, n_warm, n_regs = args
    from lego_loam_b200 import synth
    key_poses, kf, new = make_cycle_sequence((seq_id, n_kf, n_new, sensor))
    raw_c, raw_s = cpu_raw_map(key_poses, kf)
    rng = np.random.default_rng(100 + seq_id)
    inits = [synth.perturb_pose(np.asarray(x[3], np.float64), rng).astype(np.float32) for x in new]
    run, kind = cpu_cycle_factory(raw_c, raw_s, new, inits)
    for i in range(max(n_warm, 1)):
        run(i)
    _REF_BARRIER.wait()                                      # every worker has built its inputs
    t0 = time.perf_counter()
    for i in range(n_regs):
        run(i)
    return time.perf_counter() - t0, kind, int(raw_c.shape[0] + raw_s.shape[0])


def run_reference(args):
    """--impl reference: the reference's CPU path for the same mapping cycle, one registration stream per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    per_core = 1                                             # registrations per core per step
    K, W = args.steps, args.warmup
    sensor = WORKLOADS[args.workload][0]
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(cores)
    with ctx.Pool(cores, initializer=_ref_init, initargs=(barrier,)) as pool:
        res = pool.map(_ref_worker, [(c % max(args.distinct if args.distinct > 0 else 32, 1), args.key_frames, args.scans, sensor, min(W, 2), per_core * K)
                                     for c in range(cores)], chunksize=1)
    wall = max(r[0] for r in res)
    kind = res[0][1]
    total = cores * per_core * K
    value = total / wall
    line = {
        "impl": "reference", "metric": "scan-to-map registrations/s", "value": value, "unit": "registrations/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": wall / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cycle_workload_text(args.workload, args.key_frames, res[0][2], None) +
                               f"; {per_core} registration/core/step on {cores} cores"},
        "cpu_baseline": {"value": value, "unit": "registrations/s", "cores": cores, "kind": kind,
                         "sample": f"{total} registrations ({per_core}/core/step x {K} steps x {cores} cores)"},
        "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_registration_1core": wall / (per_core * K) * 1e3,
    }
    print(json.dumps(line))
    return 0


def cycle_workload_text(workload, n_kf, raw_pts, ds_pts):
    sensor = WORKLOADS[workload][0]
    t = (f"{workload}: independent {sensor.upper()} synthetic sequences, registration = one mapping cycle: local map from "
         f"{n_kf} key-frames (~{raw_pts} raw points")
    if ds_pts:
        t += f" -> ~{ds_pts}-pt voxel-DS map"
    return t + ") + 2 map voxel filters + index + downsampleCurrentScan + scan2MapOptimization"


def registration_only_arm(api, torch, local, workload, S, NB, D, n_scans, K, W):
    """Secondary arm (rank 0): the drop-in signature alone, as round 1 reported it.  The voxel-DS map is handed over
    (MO:1057-1064 is the caller's tail), a registration = downsampleCurrentScan + scan2MapOptimization (index build
    included); device-resident inputs (CUDA events) and host clouds through the C ABI (scan AND DS map cross PCIe for
    every registration: PCIe-bound)."""
    dev = torch.device("cuda", local)
    base = []
    setup_ctx = api.Context(local)
    for d in range(D):
        mc, ms, scans = make_inputs(workload, d, n_scans)
        setup_ctx.map_set_raw(mc, ms)
        base.append((setup_ctx.map_get_ds(0), setup_ctx.map_get_ds(1), scans))
    setup_ctx.close()
    seqs = []
    for s in range(S):
        mc_ds, ms_ds, scans = base[s % D]
        seqs.append({"scans": scans, "mc_ds": mc_ds, "ms_ds": ms_ds, "mc32": api.to_pcl(mc_ds), "ms32": api.to_pcl(ms_ds),
                     "scans32": [(api.to_pcl(sc.corner_last), api.to_pcl(sc.surf_last), api.to_pcl(sc.outlier_last), init)
                                 for sc, init in scans],
                     "d_mc": torch.from_numpy(mc_ds).to(dev), "d_ms": torch.from_numpy(ms_ds).to(dev),
                     "d_scans": [(torch.from_numpy(sc.corner_last).to(dev), torch.from_numpy(sc.surf_last).to(dev),
                                  torch.from_numpy(sc.outlier_last).to(dev)) for sc, init in scans]})
    torch.cuda.synchronize()
    max_map = max(max(q["mc_ds"].shape[0], q["ms_ds"].shape[0]) for q in seqs) + 1024
    max_scan = max(max(c.shape[0] for c in sc[:3]) for q in seqs for sc in q["scans32"]) + 256
    prm = api.default_params(); prm.pin_host_clouds = 1
    P = api.Batch.pack
    batches = []
    for g in [list(range(b, S, NB)) for b in range(NB)]:
        b = api.Batch(local, len(g), min(max_scan, 16384), max_map, prm)
        tabs = {"T": [np.stack([seqs[s]["scans"][i][1] for s in g]).astype(np.float32) for i in range(n_scans)]}
        for kind, dev_side in (("dev", True), ("host", False)):
            ptr = (lambda a: a.data_ptr()) if dev_side else (lambda a: a.ctypes.data)
            mk, sk = ("d_mc", "d_ms") if dev_side else ("mc32", "ms32")
            tabs[kind + "_map"] = (P([ptr(seqs[s][mk]) for s in g], [seqs[s][mk].shape[0] for s in g]),
                                   P([ptr(seqs[s][sk]) for s in g], [seqs[s][sk].shape[0] for s in g]))
            sckey = "d_scans" if dev_side else "scans32"
            tabs[kind + "_scan"] = [tuple(P([ptr(seqs[s][sckey][i][k]) for s in g], [seqs[s][sckey][i][k].shape[0] for s in g])
                                          for k in range(3)) for i in range(n_scans)]
        batches.append({"b": b, "g": g, "tabs": tabs, "stream": torch.cuda.ExternalStream(b.stream, device=local)})

    def enqueue(bt, i, kind):
        b, tabs = bt["b"], bt["tabs"]
        c, s_, o = tabs[kind + "_scan"][i % n_scans]
        b.scan_set_all(c, s_, o, dev=(kind == "dev"))
        mc, ms = tabs[kind + "_map"]
        b.map_set_ds_all(mc, ms, dev=(kind == "dev"))        # index rebuilt every registration, like the kd-trees MO:1333-1334
        b.register_async(tabs["T"][i % n_scans])

    def run_steps(kind, n):
        for i in range(n):
            for bt in batches:
                if bt.get("pending"):
                    bt["b"].result(); bt["pending"] = False
                enqueue(bt, i, kind); bt["pending"] = True
        for bt in batches:
            if bt.get("pending"):
                bt["b"].result(); bt["pending"] = False

    run_steps("dev", W)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1s = [torch.cuda.Event(enable_timing=True) for _ in batches]
    e0.record(batches[0]["stream"])
    run_steps("dev", K)
    for bt, e in zip(batches, e1s):
        e.record(bt["stream"])
    torch.cuda.synchronize()
    dev_ms = max(e0.elapsed_time(e) for e in e1s)
    barrier = threading.Barrier(NB + 1)

    def worker(k):
        torch.cuda.set_device(local)
        bt = batches[k]
        for i in range(W):
            enqueue(bt, i, "host"); bt["b"].result()
        barrier.wait()
        for i in range(K):
            enqueue(bt, i, "host"); bt["b"].result()
        barrier.wait()

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(NB)]
    for x in ths:
        x.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    host_s = time.perf_counter() - t0
    for x in ths:
        x.join()
    h2d = int(np.mean([sum(a.nbytes for a in q["scans32"][0][:3]) + q["mc32"].nbytes + q["ms32"].nbytes + 24 for q in seqs]))
    for bt in batches:
        bt["b"].close()
    return {"value_device_resident": S * K / (dev_ms * 1e-3), "e2e_host_clouds": S * K / host_s, "unit": "registrations/s",
            "sequences": S, "batches": NB, "steps": K, "h2d_bytes_per_registration": h2d,
            "map_points": int(np.mean([q["mc_ds"].shape[0] + q["ms_ds"].shape[0] for q in seqs])),
            "note": "downsampleCurrentScan + scan2MapOptimization with the voxel-DS map handed over (round-1 main arms); the "
                    "host-cloud form moves scan AND DS map over PCIe for every registration"}


def latency_arm(api, torch, local, workload, W):
    """One sequence alone on the single-registration path (one persistent kernel), L2 flushed before every registration."""
    dev = torch.device("cuda", local)
    mc, ms, scans = make_inputs(workload, 0, 2)
    setup = api.Context(local); setup.map_set_raw(mc, ms)
    mc_ds, ms_ds = setup.map_get_ds(0), setup.map_get_ds(1)
    setup.close()
    mc32, ms32 = api.to_pcl(mc_ds), api.to_pcl(ms_ds)
    scans32 = [(api.to_pcl(sc.corner_last), api.to_pcl(sc.surf_last), api.to_pcl(sc.outlier_last), init) for sc, init in scans]
    d_mc, d_ms = torch.from_numpy(mc_ds).to(dev), torch.from_numpy(ms_ds).to(dev)
    d_scans = [(torch.from_numpy(sc.corner_last).to(dev), torch.from_numpy(sc.surf_last).to(dev),
                torch.from_numpy(sc.outlier_last).to(dev)) for sc, init in scans]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lat_prm = api.default_params(); lat_prm.pin_host_clouds = 1   # one CTA per SM; host clouds DMA'd in place (long-lived members)
    lat_ctx = api.Context(local, lat_prm)
    lat_stream = torch.cuda.ExternalStream(lat_ctx.stream, device=local)
    d_T = torch.zeros(6, dtype=torch.float32, device=dev)
    d_init = [torch.from_numpy(np.asarray(init, np.float32).copy()).to(dev) for _, init in scans]
    lat, lat_host = [], []

    def step_single(i):
        c, s_, o = d_scans[i % 2]
        with torch.cuda.stream(lat_stream):
            d_T.copy_(d_init[i % 2], non_blocking=True)
        lat_ctx.scan_set_dev(c.data_ptr(), c.shape[0], s_.data_ptr(), s_.shape[0], o.data_ptr(), o.shape[0])
        lat_ctx.downsample_current_scan(want_counts=False)
        lat_ctx.map_set_ds_dev(d_mc.data_ptr(), d_mc.shape[0], d_ms.data_ptr(), d_ms.shape[0])
        lat_ctx.s2m_optimize_dev(d_T.data_ptr())

    with torch.cuda.stream(lat_stream):
        for i in range(W + 20):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(lat_stream); step_single(i); b.record(lat_stream)
            lat_stream.synchronize()
            if i >= W:
                lat.append(a.elapsed_time(b))
        for i in range(W + 20):                              # the same through the C ABI with host clouds, wall clock
            c, s_, o, init = scans32[i % 2]
            flush.zero_(); lat_stream.synchronize()
            t0 = time.perf_counter()
            lat_ctx.scan_set_pcl(c, s_, o); lat_ctx.downsample_current_scan(want_counts=False)
            lat_ctx.map_set_ds_pcl(mc32, ms32); lat_ctx.s2m_optimize(init)
            if i >= W:
                lat_host.append((time.perf_counter() - t0) * 1e3)
    del flush
    lat_ctx.close()
    return {"ms_per_scan_device": float(np.median(lat)), "ms_per_scan_device_max": float(np.max(lat)),
            "ms_per_scan_e2e_host": float(np.median(lat_host)), "ms_per_scan_e2e_host_max": float(np.max(lat_host)),
            "map_points": int(mc_ds.shape[0] + ms_ds.shape[0]),
            "note": "one sequence alone on the single-registration path (one persistent kernel, one CTA per SM), DS map "
                    "handed over, L2 flushed before each registration; e2e_host = host PCL clouds in (page-locked once, DMA "
                    "in place), pose out, wall clock"}


def odometry_arm(api, local, reps, cpu_sample):
    """Secondary arm for the featureAssociation rows of the path: updateTransformation (FA:1666-1695:
    both <= 25-iteration LM loops with their correspondence searches) for one VLP-16 sweep pair, device time and wall
    time through the C ABI with host clouds, beside the reference's own function on one core."""
    from lego_loam_b200 import synth
    w = synth.make_world(synth.SEED0)
    od = synth.make_odometry_pair(w, synth.VLP16, np.array([0, 1.0, 0, 3.0, 0, -4.0]),
                                  np.array([0.002, 0.015, -0.001, 0.01, 0.005, -0.15]), seed=1)
    c = api.Context(local)
    dev_ms, wall_ms = [], []
    for i in range(reps + 3):
        t0 = time.perf_counter()
        c.odom_set_last(od.corner_last, od.surf_last)            # FA:1615-1619 / FA:1774-1788 (index of the last sweep)
        c.odom_set_features(od.corner_sharp, od.surf_flat)
        T, s0, s1 = c.odom_optimize(np.zeros(6, np.float32))
        if i >= 3:
            wall_ms.append((time.perf_counter() - t0) * 1e3); dev_ms.append(s0.device_ms)
    c.close()
    # throughput form: 64 sweep pairs per launch on the batched engine (one persistent CTA per pair)
    NBO = 64
    b = api.Batch(local, NBO, 8192, 64)
    pcl = tuple(api.to_pcl(x) for x in (od.corner_last, od.surf_last, od.corner_sharp, od.surf_flat))
    bt_wall = []
    for i in range(reps + 3):
        t0 = time.perf_counter()
        for s in range(NBO):
            b._ck(api.lib().llb_batch_odom_set(b._h, s, api._vp(pcl[0]), pcl[0].shape[0], api._vp(pcl[1]), pcl[1].shape[0],
                                               api._vp(pcl[2]), pcl[2].shape[0], api._vp(pcl[3]), pcl[3].shape[0]))
        Tb, b0, b1 = b.odom_optimize(np.zeros((NBO, 6), np.float32))
        if i >= 3:
            bt_wall.append((time.perf_counter() - t0) * 1e3)
    batch_dev_ms = float(b0[0].device_ms)
    batch_equal = bool(np.array_equal(Tb[0], np.asarray(T, np.float32)) and np.array_equal(Tb[NBO - 1], np.asarray(T, np.float32)))
    b.close()
    kind, fa = "port", None
    try:
        from oracle import ref_harness
        if ref_harness.available():
            kind = "reference"; fa = ref_harness.FeatureAssociation()
    except Exception:
        pass
    if fa is None:
        import oracle
        oracle.set_trig_mode(0); fa = oracle.FeatureAssociation()

    def cpu_once():
        fa.set_last(od.corner_last, od.surf_last, True)
        fa.set_features(od.corner_sharp, od.surf_flat)
        fa.transformCur = np.zeros(6, np.float32)
        fa.updateTransformation()
        return fa.transformCur
    cpu_once()
    t0 = time.perf_counter()
    for _ in range(cpu_sample):
        Tc = cpu_once()
    cpu_ms = (time.perf_counter() - t0) / cpu_sample * 1e3
    return {"ms_per_scan_device": float(np.median(dev_ms)), "ms_per_scan_e2e_host": float(np.median(wall_ms)),
            "iterations": [int(s0.iterations), int(s1.iterations)],
            "features": {"sharp": int(od.corner_sharp.shape[0]), "flat": int(od.surf_flat.shape[0]),
                         "corner_last": int(od.corner_last.shape[0]), "surf_last": int(od.surf_last.shape[0])},
            "cpu_1core": {"ms_per_scan": cpu_ms, "kind": kind, "sample": f"{cpu_sample} calls"},
            "pose_max_abs_diff_vs_cpu": float(np.max(np.abs(np.asarray(T) - np.asarray(Tc)))),
            "batched": {"pairs_per_launch": NBO, "ms_per_launch_device": batch_dev_ms,
                        "pairs_per_s_e2e_host": NBO / (float(np.median(bt_wall)) * 1e-3),
                        "equal_to_single": batch_equal},
            "note": "updateTransformation (FA:1666-1695) of one VLP-16 sweep pair: one persistent 8-CTA cluster on the device "
                    "(redundant iteration loop, correspondence search split by feature); batched: one CTA per pair"}


def feature_extraction_arm(api, local, reps, cpu_sample):
    """Secondary arm for the next row of the path: adjustDistortion + calculateSmoothness +
    markOccludedPoints + extractFeatures (FA:491-784) of one segmented VLP-16 sweep: device time, wall time through the
    C ABI with host buffers (segmented cloud + cloud_info in, four feature clouds out), and the reference's own
    functions on one core."""
    from lego_loam_b200 import synth
    w = synth.make_world(synth.SEED0)
    sws = [synth.make_segmented_sweep(w, synth.VLP16, [0, 0.05 + 0.01 * k, 0, 3 + 0.4 * k, 0, 5], 11 + k) for k in range(4)]
    c = api.Context(local); c.features_init(16, 1800)
    dev_ms, wall_ms = [], []
    for i in range(reps + 3):
        t0 = time.perf_counter()
        counts, ms = c.features_extract(sws[i % 4])
        if i >= 3:
            wall_ms.append((time.perf_counter() - t0) * 1e3); dev_ms.append(ms)
    got = [c.features_get(k) for k in range(4)]
    last = sws[(reps + 2) % 4]
    c.close()
    kind, fa = "port", None
    try:
        from oracle import ref_harness
        if ref_harness.available():
            kind = "reference"; fa = ref_harness.FeatureAssociation()
    except Exception:
        pass
    if fa is None:
        import oracle
        fe = oracle.FeatureExtraction(16, 1800)
        run = lambda sw: fe.extract(sw)[:4]
    else:
        def run(sw):
            fa.set_segmented(sw); fa.extract_features()
            return [fa.feature_cloud(k) for k in range(4)]
    # same sweep sequence through ONE CPU object (state survives between sweeps, as on the device)
    cpu_t = []
    for i in range(reps + 3):
        t0 = time.perf_counter(); ref = run(sws[i % 4]); cpu_t.append((time.perf_counter() - t0) * 1e3)
    same_xyz = all(g.shape == r.shape and np.array_equal(g[:, :3].view(np.uint32), np.asarray(r)[:, :3].view(np.uint32))
                   for g, r in zip(got, ref))
    dint = max(float(np.max(np.abs(g[:, 3] - np.asarray(r)[:, 3]))) if g.size and g.shape == r.shape else 0.0 for g, r in zip(got, ref))
    # FA side of a node cycle, device-resident: extractFeatures -> updateTransformation -> publishCloudsLast; only the
    # segmented cloud goes up, only the pose comes back
    cyc = api.Context(local); cyc.features_init(16, 1800)
    cyc_wall, T = [], np.zeros(6, np.float32)
    for i in range(reps + 3):
        t0 = time.perf_counter()
        cyc.features_extract(sws[i % 4])
        if i > 0:
            cyc.features_to_odometry()
            T, s0, s1 = cyc.odom_optimize(np.zeros(6, np.float32))
        cyc.features_publish_last(T)
        cyc.synchronize()
        if i >= 3:
            cyc_wall.append((time.perf_counter() - t0) * 1e3)
    cyc.close()
    # throughput form: 64 sequences, one sweep each per step, on the batched engine (one set of five launches per step)
    NBF = 64

    class Held:          # a sweep with its 32 B-stride cloud built once (the caller's PCL cloud)
        def __init__(self, sw):
            self.__dict__.update(sw.__dict__); self.cloud32 = api.to_pcl(sw.cloud)
    held = [Held(s) for s in sws]
    fb = api.Batch(local, NBF, 4096, 4096); fb.features_init(16, 1800)
    packs = [fb.features_pack([held[(s + i) % 4] for s in range(NBF)]) for i in range(2)]
    fb_dev, fb_wall = [], []
    for i in range(reps + 3):
        t0 = time.perf_counter(); bcounts, bms = fb.features_extract(packs[i % 2]); t1 = time.perf_counter()
        if i >= 3:
            fb_dev.append(bms); fb_wall.append((t1 - t0) * 1e3)
    fb.close()
    batched = {"sweeps_per_step": NBF, "ms_per_step_device": float(np.median(fb_dev)),
               "sweeps_per_s_e2e_host": NBF / (float(np.median(fb_wall)) * 1e-3),
               "note": "segmented clouds + cloud_info of all slots in from host buffers, four feature clouds per slot out"}
    cpu_cyc = None
    if fa is not None:
        fa2 = ref_harness.FeatureAssociation()
        t_cpu = []
        for i in range(reps + 3):
            t0 = time.perf_counter()
            fa2.set_segmented(sws[i % 4]); fa2.extract_features()
            fa2.transformCur = np.zeros(6, np.float32)
            if i > 0:
                fa2.updateTransformation()
            fa2.publishCloudsLast()
            t_cpu.append((time.perf_counter() - t0) * 1e3)
        cpu_cyc = {"ms_per_sweep": float(np.median(t_cpu[3:])), "kind": "reference",
                   "pose_max_abs_diff": float(np.max(np.abs(np.asarray(fa2.transformCur) - np.asarray(T))))}
    return {"ms_per_sweep_device": float(np.median(dev_ms)), "ms_per_sweep_e2e_host": float(np.median(wall_ms)),
            "points": int(last.cloud.shape[0]), "counts": [int(x) for x in counts],
            "cpu_1core": {"ms_per_sweep": float(np.median(cpu_t[3:])), "kind": kind,
                          "sample": f"{reps} sweeps (includes the harness copies of the clouds in and out)"},
            "selection_and_xyz_identical_to_cpu": bool(same_xyz), "intensity_max_abs_diff_vs_cpu": dint,
            "batched": batched,
            "fa_cycle": {"ms_per_sweep_e2e_host": float(np.median(cyc_wall)), "cpu_1core": cpu_cyc,
                         "what": "extractFeatures + updateTransformation + publishCloudsLast (TransformToEnd, last clouds, index) "
                                 "per sweep; device: one upload (segmented cloud + cloud_info), pose back"},
            "note": "5 launches per sweep: per-point kernels, one CTA per ring (sector sorts = libstdc++ std::sort move "
                    "for move, greedy picks 32 candidates per step), per-ring VoxelGrid(0.2), concatenation"}


def sharded_arm(api, torch, dist, local, rank, world, workload="vls128_2m"):
    """BASELINE config 4 (only with WORLD_SIZE > 1): ONE mapping cycle of a VLS-128 sweep against a ~2M-point voxel-DS map
    (8M raw points, device-resident on every rank as replicated key-frame stores leave them).
    map_sharded (preferred form): every rank voxel-filters and indexes only its slab of the raw map (on the
    lattice of the whole map) and takes the queries inside its slab, so the map-side work AND the kNN + fit work divide by
    the number of ranks; the 28 fp64 sums of every LM iteration are exchanged by P2P stores over NVLink fused into the
    persistent kernel.  query_sharded (round 1): DS map + index replicated, queries dealt round-robin, exchange fused or
    by an NCCL all-reduce driven from the host.  Poses must be bit-identical across ranks and equal the single-GPU pose."""
    from lego_loam_b200 import multi_gpu
    mc, ms, scans = make_inputs(workload, 0, 1)              # identical inputs on every rank (same seeds)
    sc, init = scans[0]
    dev = torch.device("cuda", local)
    mc_d = torch.as_tensor(np.ascontiguousarray(mc, np.float32), device=dev)
    ms_d = torch.as_tensor(np.ascontiguousarray(ms, np.float32), device=dev)
    ptrs = (mc_d.data_ptr(), int(mc.shape[0]), ms_d.data_ptr(), int(ms.shape[0]))
    torch.cuda.synchronize()
    ctx = api.Context(local)

    def timed(fn, reps, skip=1):
        out = []
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter(); r = fn(); ctx.synchronize()
            out.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(out[skip:])), r

    # ---- one GPU alone: map voxel filters + index over the whole raw map, then the registration
    map_single_ms, _ = timed(lambda: ctx.map_set_raw_dev(*ptrs), 4)
    n_ds = [int(ctx.map_get_ds(0).shape[0]), int(ctx.map_get_ds(1).shape[0])]
    ctx.scan_set(sc.corner_last, sc.surf_last, sc.outlier_last)
    counts = ctx.downsample_current_scan()
    single = []
    for _ in range(5):
        T_single, st = ctx.s2m_optimize(init); single.append(st.device_ms)
    # ---- query-sharded over the replicated DS map (round-1 form): NCCL-driven and fused exchange
    nccl_ms, (T_nccl, it_nccl) = timed(lambda: multi_gpu.sharded_scan2map(ctx, init, rank, world), 4)
    multi_gpu.setup_fused_exchange(ctx, rank, world)
    fused_dev = []
    fused_wall, (T_fused, st_f) = timed(lambda: (lambda r: (fused_dev.append(r[1].device_ms), r)[1])(multi_gpu.sharded_scan2map_fused(ctx, init)), 8, 2)
    # ---- map-sharded: slab of the raw map per rank (filters + index divide by the number of ranks), fused exchange
    infos = []
    map_shard_ms, _ = timed(lambda: infos.append(multi_gpu.set_sharded_map(ctx, None, None, rank, world, device_ptrs=ptrs)), 4)
    info = infos[-1]
    ctx.scan_set(sc.corner_last, sc.surf_last, sc.outlier_last); ctx.downsample_current_scan()
    slab_dev = []
    slab_wall, (T_slab, st_s) = timed(lambda: (lambda r: (slab_dev.append(r[1].device_ms), r)[1])(multi_gpu.sharded_scan2map_fused(ctx, init)), 8, 2)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(fused=np.asarray(T_fused, np.float32).tobytes(), slab=np.asarray(T_slab, np.float32).tobytes(),
                                          fused_dev=float(np.median(fused_dev[2:])), slab_dev=float(np.median(slab_dev[2:])),
                                          map_shard_ms=map_shard_ms, map_single_ms=map_single_ms,
                                          raw_kept=int(info.raw_kept[0] + info.raw_kept[1]), ds_local=int(info.ds_local[0] + info.ds_local[1])))
    g0 = gathered[0]
    t_single = np.asarray(T_single, np.float32).tobytes()
    map_sh = max(g["map_shard_ms"] for g in gathered); slab_reg = max(g["slab_dev"] for g in gathered)
    out = {"workload": workload, "world": world, "queries": int(counts[0] + counts[3]), "map_points": n_ds,
           "raw_map_points": int(mc.shape[0] + ms.shape[0]), "iterations": int(st.iterations),
           "single_gpu": {"map_voxel_index_ms": map_single_ms, "registration_device_ms": float(np.median(single[1:])),
                          "cycle_ms": map_single_ms + float(np.median(single[1:]))},
           "map_sharded": {"map_voxel_index_ms_max_over_ranks": map_sh, "registration_device_ms_max_over_ranks": slab_reg,
                           "registration_wall_ms": slab_wall, "cycle_ms": map_sh + slab_reg,
                           "raw_points_per_rank": [g["raw_kept"] for g in gathered], "ds_points_per_rank": [g["ds_local"] for g in gathered],
                           "axis": int(info.axis), "iterations": int(st_s.iterations),
                           "pose_bit_identical_across_ranks": all(g["slab"] == g0["slab"] for g in gathered),
                           "pose_equals_single_gpu": g0["slab"] == t_single,
                           "note": "map_voxel_index_ms = wall clock of the call incl. the slab planning (one small D2H + host sort) and "
                                   "the all-reduce of the owned DS counts; raw map device-resident on every rank"},
           "query_sharded": {"fused_device_ms_max_over_ranks": float(max(g["fused_dev"] for g in gathered)), "fused_wall_ms": fused_wall,
                             "nccl_wall_ms": nccl_ms, "map_voxel_index_ms": max(g["map_single_ms"] for g in gathered),
                             "pose_bit_identical_across_ranks": all(g["fused"] == g0["fused"] for g in gathered),
                             "pose_equals_single_gpu": g0["fused"] == t_single,
                             "fused_equals_nccl": bool(np.array_equal(np.asarray(T_fused, np.float32), np.asarray(T_nccl, np.float32)))},
           "bytes_per_iteration_per_rank": 28 * 8 * (world - 1),
           "speedup_cycle_vs_single_gpu": (map_single_ms + float(np.median(single[1:]))) / (map_sh + slab_reg)}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vlp16_100k", choices=sorted(WORKLOADS))
    ap.add_argument("--seqs", type=int, default=64, help="independent sequences registered per step per GPU")
    ap.add_argument("--batches", type=int, default=2, help="batch objects (streams) the sequences are split over")
    ap.add_argument("--distinct", type=int, default=0, help="distinct synthetic sequences generated per GPU (slots beyond reuse them with their own initial guesses); "
                    "0 = 32 on one GPU, 16 per GPU under torchrun (the ranks share the host cores that ray-cast them: untimed set-up)")
    ap.add_argument("--scans", type=int, default=8, help="distinct new sweeps per sequence rotated through the steps")
    ap.add_argument("--cpu-sample", type=int, default=8, help="registrations timed for cpu_baseline")
    ap.add_argument("--key-frames", type=int, default=100, help="resident key-frames the local map of a sequence is assembled from")
    ap.add_argument("--sharded", type=int, default=1, help="1: with WORLD_SIZE > 1 also run the config-4 arm (one VLS-128 registration sharded over the ranks)")
    ap.add_argument("--secondary", type=int, default=1, help="1: also run the secondary arms on rank 0 (registration_only, latency, odometry, feature_extraction)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from lego_loam_b200 import api, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W, S = args.steps, max(args.warmup, 3), args.seqs
    NB = max(1, min(args.batches, S))
    KF, NS = args.key_frames, max(args.scans, 1)
    dev = torch.device("cuda", local)
    sensor = WORKLOADS[args.workload][0]

    # ---------------- D distinct sequences per rank (ray-cast on the host, all cores, untimed); slot s replays
    # sequence s % D with its own device copies of the sweeps and its own initial guesses
    D = max(1, min(args.distinct if args.distinct > 0 else (32 if world == 1 else 16), S))
    seq = make_cycle_sequences([1000 * rank + d for d in range(D)], KF, NS, sensor)
    _log("sequences generated")
    host32 = [[tuple(api.to_pcl(x) for x in sw[:3]) for sw in seq[d][2]] for d in range(D)]      # the caller's PCL clouds
    kf32 = [[tuple(api.to_pcl(x) for x in f) for f in seq[d][1]] for d in range(D)]
    max_scan = max(max(x.shape[0] for f in seq[d][1] + [sw[:3] for sw in seq[d][2]] for x in f) for d in range(D)) + 256
    prm = api.default_params(); prm.pin_host_clouds = 0      # host sweeps are packed into one pinned block per step: ONE copy
    P = api.Batch.pack
    per = S // NB
    ids = np.arange(KF, dtype=np.int32)
    batches = []
    for bi in range(NB):
        g = list(range(bi, S, NB))[:per]
        b = api.Batch(local, per, min(max_scan, 16384), 4096, prm)
        b.enable_keyframes(int(KF * 9000 * (2.0 if sensor != "vlp16" else 1.0)), KF)
        sd = [s % D for s in g]
        dummy = api.to_pcl(np.zeros((16, 4), np.float32))    # placeholder map while the key-frames are collected
        for s in range(per):
            b.map_set_ds_pcl(s, dummy, dummy)
        for k in range(KF):                                  # saveKeyFramesAndFactor's cloud part, MO:1443-1453
            tab = tuple(P([kf32[d][k][j].ctypes.data for d in sd], [kf32[d][k][j].shape[0] for d in sd]) for j in range(3))
            b.scan_set_all(*tab, dev=False)
            b.register(np.zeros((per, 6), np.float32))       # (skipped by the guard MO:1331: only the DS clouds matter)
            for s in range(per):
                b.keyframe_add(s)
        rng = np.random.default_rng(100 + 1000 * rank + bi)
        init = [np.stack([synth.perturb_pose(np.asarray(seq[d][2][i][3], np.float64), rng) for d in sd]).astype(np.float32)
                for i in range(NS)]
        d_sw = [[tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in seq[d][2][i][:3]) for i in range(NS)] for d in sd]
        tabs = {"host": [tuple(P([host32[d][i][j].ctypes.data for d in sd], [host32[d][i][j].shape[0] for d in sd]) for j in range(3))
                         for i in range(NS)],
                "dev": [tuple(P([d_sw[s][i][j].data_ptr() for s in range(per)], [d_sw[s][i][j].shape[0] for s in range(per)])
                              for j in range(3)) for i in range(NS)]}
        batches.append({"b": b, "g": g, "sd": sd, "tabs": tabs, "init": init, "keep": d_sw, "dummy": dummy,
                        "asm": api.Batch.pack_assemble([ids] * per, [seq[d][0] for d in sd]),
                        "stream": torch.cuda.ExternalStream(b.stream, device=local)})
    # (kf32 stays alive: with pin_host_clouds the library page-locks the caller's buffers in place, and freed-but-registered
    # memory handed out again by malloc breaks later copies)
    torch.cuda.synchronize()
    _log(f"set-up done: {D} sequences x ({KF} key-frames + {NS} sweeps), {NB} batches x {per} slots")

    def enqueue(bt, i, kind):
        """one step of one batch: the new sweep of every slot, its local map from the resident key-frames, registration"""
        b = bt["b"]
        b.scan_set_all(*bt["tabs"][kind][i % NS], dev=(kind == "dev"))
        b.map_assemble_all(bt["asm"])                        # resident key-frames -> raw map -> 2 voxel filters -> index
        b.register_async(bt["init"][i % NS])

    def run_steps(kind, n):
        """n steps of every batch, software-pipelined from one host thread: while batch A computes, B is prepared"""
        last = None
        for i in range(n):
            for bt in batches:
                if bt.get("pending"):
                    last = bt["b"].result(); bt["pending"] = False
                enqueue(bt, i, kind); bt["pending"] = True
        for bt in batches:
            if bt.get("pending"):
                last = bt["b"].result(); bt["pending"] = False
        return last

    run_steps("dev", W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk_samples, stop_evt = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop_evt, clk_samples, local)); th.start()
    l0 = sum(bt["b"].launch_count() for bt in batches)
    e_start = torch.cuda.Event(enable_timing=True)
    e_ends = [torch.cuda.Event(enable_timing=True) for _ in batches]
    # profilers attach to the timed region only (ncu --profile-from-start off): the set-up is thousands of launches
    cu_prof = None
    if os.environ.get("LLB_BENCH_CUPROF"):
        import ctypes
        cu_prof = ctypes.CDLL("libcuda.so.1"); cu_prof.cuProfilerStart()
    t_wall0 = time.perf_counter()
    e_start.record(batches[0]["stream"])
    run_steps("dev", K)
    for bt, e in zip(batches, e_ends):
        e.record(bt["stream"])
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    if cu_prof is not None:
        cu_prof.cuProfilerStop()
    launches = sum(bt["b"].launch_count() for bt in batches) - l0
    if world > 1:
        dist.barrier()
    total_ms = max(e_start.elapsed_time(e) for e in e_ends)
    _log(f"value arm done: {total_ms / K:.3f} ms per step")

    # ---------------- stage times and the roofline of the registration kernel, with CUDA events on the batch stream
    # during extra steps after the timed region (per-stage events perturb the pipelining)
    b0 = batches[0]
    b0["b"].set_profile(True)
    prof_acc, geo, qi_acc = {}, None, 0
    NPROF = 5
    for i in range(NPROF):
        enqueue(b0, i, "dev"); Tp, stp = b0["b"].result()
        pr, geo = b0["b"].get_profile()
        for k, v in pr.items():
            prof_acc[k] = prof_acc.get(k, 0.0) + v / NPROF
        qi_acc += sum((x.n_corner_ds + x.n_surf_ds) * x.iterations for x in stp) / NPROF
    b0["b"].set_profile(False)
    raw_n = [(int(b0["b"].map_get(s, 0).shape[0]), int(b0["b"].map_get(s, 1).shape[0])) for s in range(min(per, D))]
    ds_n = [(int(b0["b"].map_get(s, 2).shape[0]), int(b0["b"].map_get(s, 3).shape[0])) for s in range(min(per, D))]
    raw_c0, raw_s0 = b0["b"].map_get(0, 0), b0["b"].map_get(0, 1)

    # ---------------- end-to-end arm on every rank: host sweeps through the C ABI, one host thread per batch (the H2D
    # and host work of one batch overlap the kernels of the other)
    barrier = threading.Barrier(NB + 1)
    last = {}

    def worker(k):
        torch.cuda.set_device(local)
        bt = batches[k]
        for i in range(W):
            enqueue(bt, i, "host"); bt["b"].result()
        barrier.wait()
        # the next sweeps are handed over while the current registrations run (the library packs them into one pinned block
        # and sends it with one copy on its own stream): host packing and PCIe overlap the kernels of the step in flight
        b = bt["b"]
        b.scan_set_all(*bt["tabs"]["host"][0], dev=False)
        for i in range(K):
            b.map_assemble_all(bt["asm"])
            b.register_async(bt["init"][i % NS])
            if i + 1 < K:
                b.scan_set_all(*bt["tabs"]["host"][(i + 1) % NS], dev=False)
            res = b.result()
        if k == 0:
            last["T"], last["st"] = res
        barrier.wait()

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(NB)]
    for x in ths:
        x.start()
    barrier.wait()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    barrier.wait()
    e2e_s = time.perf_counter() - t0
    for x in ths:
        x.join()
    if world > 1:
        dist.barrier()
    stop_evt.set(); th.join()
    _log(f"e2e arm done: {e2e_s / K * 1e3:.3f} ms per step")
    for bt in batches[1:]:
        bt["b"].close()

    sec = {}
    if args.secondary and rank == 0:
        for name, fn in (("registration_only", lambda: registration_only_arm(api, torch, local, args.workload, S, NB, min(D, 8), 2, min(K, 40), W)),
                         ("latency", lambda: latency_arm(api, torch, local, args.workload, W)),
                         ("odometry", lambda: odometry_arm(api, local, 20, 10)),
                         ("feature_extraction", lambda: feature_extraction_arm(api, local, 20, 10))):
            try:
                sec[name] = fn()
            except Exception as e:                            # secondary arms never hide the main line
                sec[name] = {"error": repr(e)}
            _log(f"secondary arm {name} done")
    shard = None
    if world > 1 and args.sharded:
        try:
            shard = sharded_arm(api, torch, dist, local, rank, world)
        except Exception as e:
            shard = {"error": repr(e)}
        _log("sharded arm done")
    if world > 1:
        dist.barrier()

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        ms_per_step = total_ms / K
        value = world * S * K / (total_ms / 1e3)
        h2d_step = int(sum(sum(a.nbytes for a in host32[d][0]) for bt in batches for d in bt["sd"]) + S * (KF * 28 + 24))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"
        # registration kernel: ONE launch per step and batch covers every LM iteration of its slots
        kern_ms = prof_acc["fit"] + prof_acc["knn"]
        alg_bytes = ALG_BYTES_PER_QUERY * qi_acc
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:                                                 # dram bytes per launch from the committed ncu capture (same workload)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if args.workload == tj.get("workload"):
                traffic = int(tj["dram_bytes_per_launch"] * per / tj["slots"])
        except Exception:
            pass
        raw_tot = float(np.mean([a + b for a, b in raw_n])); ds_tot = float(np.mean([a + b for a, b in ds_n]))
        stages = {
            "index_build": {"algorithmic_bytes": 32.0 * ds_tot * per, "ms": prof_acc["index_build"],
                            "what": "read + write re-ordered float4 of every DS map point (SURVEY 8(d)), 2 maps per slot"},
            "map_assembly_and_voxel": {"algorithmic_bytes": (32.0 * raw_tot + 16.0 * (raw_tot + ds_tot)) * per, "ms": prof_acc["unpack"],
                                       "what": "transform + concatenate the key-frame clouds (32 B per raw point), then 16 B per raw "
                                               "point in and per DS point out of the two voxel filters"},
        }
        for v in stages.values():
            v["achieved_gbs"] = v["algorithmic_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None
            v["frac"] = v["achieved_gbs"] / peak if v["achieved_gbs"] else None
        # bounded CPU sample on this box's host cores, 1 core (the reference's statements for the same cycle, raw map of
        # slot 0 as the device assembled it); also the pose check of the e2e result against the CPU
        sd0 = b0["sd"]
        n_cpu = max(args.cpu_sample, 1)
        run_cpu, kind = cpu_cycle_factory(raw_c0, raw_s0, seq[sd0[0]][2], [b0["init"][i][0] for i in range(NS)])
        run_cpu(0)
        t0 = time.perf_counter()
        for i in range(n_cpu):
            run_cpu(i)
        cpu_s = time.perf_counter() - t0
        pose_diff = None
        if "T" in last:                                      # the distinct sequences of batch 0 against the CPU reference
            pose_diff = 0.0
            for j in range(min(per, D, 4)):
                rc, _ = cpu_cycle_factory(b0["b"].map_get(j, 0), b0["b"].map_get(j, 1), seq[sd0[j]][2], [b0["init"][i][j] for i in range(NS)])
                pose_diff = max(pose_diff, float(np.max(np.abs(last["T"][j] - rc(K - 1)))))
        line = {
            "metric": "scan-to-map registrations/s", "value": value, "unit": "registrations/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cycle_workload_text(args.workload, KF, int(raw_tot), int(ds_tot)) +
                                   f"; step = one registration for each of {S} sequences per GPU, batched engine "
                                   f"({NB} batches x {per} slots), key-frame clouds device-resident",
                       "sequences_per_gpu": S, "batches": NB, "distinct_sequences": D, "sweeps_per_sequence": NS, "key_frames": KF,
                       "registrations_per_step": S * world,
                       "queries_per_registration": int(np.mean([x.n_corner_ds + x.n_surf_ds for x in stp])),
                       "raw_map_points": int(raw_tot), "map_points": int(ds_tot),
                       "l2": f"no flush: every slot has its own key-frames, raw map, DS map, index and sweeps, "
                             f"~{S * (raw_tot * 48 + ds_tot * 64) / 1e6:.0f} MB touched per step (> 126 MB L2); the latency arm flushes "
                             f"L2 (256 MiB write) before every registration",
                       "timing": "value: CUDA events, first start to last end over the batch streams (host gaps included), max "
                                 "over ranks; e2e: wall clock on every rank, max over ranks",
                       "e2e_host_threads": NB, "pin_host_clouds": 0},
            "e2e": {"value": world * S * K / e2e_s, "unit": "registrations/s",
                    "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": int(72 * S),
                    "ms_per_step": e2e_s / K * 1e3, "h2d_gbs_per_rank": h2d_step / (e2e_s / K) / 1e9,
                    "note": "host sweeps (pcl::PointXYZI, 32 B per point) in, pose + stats out through the C ABI on every rank; "
                            "the key-frame clouds are device-resident (each crossed PCIe once, as the sweep it was)"},
            "gpu_launches": int(launches),
            "clocks": summarize_clocks(clk_samples),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_source": "ncu --set full capture of the kernel, profiles/r02_traffic.json (same workload)",
                         "kernel": "batch_lm_kernel (kNN + line/plane fits + Jacobian rows + fp64 products + LM steps of ALL "
                                   "iterations of a batch's slots: one launch per step and batch)",
                         "peak_source": peak_src,
                         "ms_per_launch": kern_ms, "alg_bytes_per_launch": alg_bytes,
                         "slots_per_launch": per, "query_iterations_per_launch": qi_acc,
                         "stage_ms_per_step": prof_acc, "geometry": geo,
                         "note": "latency/issue-bound, not bandwidth-bound: ~5k warp instructions per 32 query-iterations for "
                                 "3 kB of algorithmic bytes (DESIGN.md section 3, profiles/r02_knnfit.md)"},
            "roofline_stages": stages,
            "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "registrations/s", "cores": 1, "kind": kind,
                             "sample": f"{n_cpu} mapping cycles of sequence 0 (same raw map, same sweeps), 1 core",
                             "ms_per_registration": cpu_s / n_cpu * 1e3},
            "wall_s_timed_region": t_wall,
            "last_stats": last["st"][0].as_dict() if "st" in last else None,
            "pose_check_max_abs_diff_vs_cpu": pose_diff,
        }
        line.update(sec)
        if shard is not None:
            line["sharded"] = shard
        print(json.dumps(line))
    b0["b"].close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
