#!/usr/bin/env python
"""profiles/r02_summary.md from the bench lines of the round (usage: r02_summary.py bench_n1.json [bench_n2.json])."""
import json
import sys


def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def main():
    d = last_json(sys.argv[1])
    r = d["roofline"]; st = r["stage_ms_per_step"]; rs = d["roofline_stages"]
    out = ["# round 2: numbers of the committed state (one B200, `python bench.py`; line in `r02_bench_line.json`)\n",
           "| what | value |", "|---|---|",
           f"| **value** (registrations/s, sweeps already in HBM; a registration = one mapping cycle) | **{d['value']:.0f}** ({d['ms_per_step']:.3f} ms per step of {d['config']['registrations_per_step']}) |",
           f"| **e2e** (host sweeps in through the C ABI, pose out) | **{d['e2e']['value']:.0f}** ({d['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB H2D per step) |",
           f"| cpu_baseline (the reference's statements for the same cycle, {d['cpu_baseline']['cores']} core) | {d['cpu_baseline']['value']:.1f} registrations/s ({d['cpu_baseline']['ms_per_registration']:.1f} ms each) |",
           f"| launches in the timed region | {d['gpu_launches']} ({d['gpu_launches'] / d['steps'] / d['config']['batches']:.0f} per batch step, replayed as one CUDA graph) |",
           f"| clocks | {d['clocks']['sm_mhz']} / {d['clocks']['sm_max_mhz']} MHz, reasons {d['clocks']['reasons']} |",
           f"| roofline: `batch_lm_kernel` | {r['achieved']:.1f} GB/s of {r['peak']} = **{r['frac']:.4f}**; {r['ms_per_launch'] * 1e3:.0f} µs per launch for {r['query_iterations_per_launch']:.0f} query-iterations (96 B each); DRAM traffic {r['traffic']} B per launch |",
           f"| stage: map assembly + 2 voxel filters (32 slots) | {st['unpack']:.3f} ms, {rs['map_assembly_and_voxel']['achieved_gbs']:.0f} GB/s = {rs['map_assembly_and_voxel']['frac']:.3f} of HBM on {rs['map_assembly_and_voxel']['algorithmic_bytes'] / 1e6:.0f} MB algorithmic |",
           f"| stage: index build (64 maps) | {st['index_build']:.3f} ms, {rs['index_build']['achieved_gbs']:.0f} GB/s = {rs['index_build']['frac']:.3f} |",
           f"| stage: registration kernel / LM prepare + collect | {st['fit']:.3f} / {st['lm_step']:.3f} ms |"]
    if "registration_only" in d:
        ro = d["registration_only"]
        out.append(f"| registration_only (round-1 arms: DS map handed over) | {ro['value_device_resident']:.0f} device-resident / {ro['e2e_host_clouds']:.0f} with scan AND map over PCIe |")
    if "latency" in d:
        la = d["latency"]
        out.append(f"| latency, one registration alone (L2 flushed) | {la['ms_per_scan_device']:.3f} ms device / {la['ms_per_scan_e2e_host']:.3f} ms from host clouds |")
    if "odometry" in d:
        od = d["odometry"]
        out.append(f"| odometry (`updateTransformation`) | {od['ms_per_scan_device']:.3f} ms device / {od['ms_per_scan_e2e_host']:.3f} ms host; reference {od['cpu_1core']['ms_per_scan']:.2f} ms on one core; batched {od['batched']['pairs_per_s_e2e_host']:.0f} pairs/s |")
    if "feature_extraction" in d:
        fe = d["feature_extraction"]
        out.append(f"| feature extraction | {fe['ms_per_sweep_device']:.3f} ms device / {fe['ms_per_sweep_e2e_host']:.3f} ms host; reference {fe['cpu_1core']['ms_per_sweep']:.2f} ms; FA cycle {fe['fa_cycle']['ms_per_sweep_e2e_host']:.3f} ms vs {fe['fa_cycle']['cpu_1core']['ms_per_sweep']:.2f} ms |")
    if len(sys.argv) > 2:
        d2 = last_json(sys.argv[2]); s = d2["sharded"]
        out += ["", f"## N = {d2['n_gpus']} (`r02_bench_n2_line.json`)\n", "| what | value |", "|---|---|",
                f"| value / e2e (replicas, weak scaling) | {d2['value']:.0f} / {d2['e2e']['value']:.0f} registrations/s |",
                f"| config 4, one GPU: map voxel filters + index / registration / cycle | {s['single_gpu']['map_voxel_index_ms']:.3f} / {s['single_gpu']['registration_device_ms']:.3f} / {s['single_gpu']['cycle_ms']:.3f} ms ({s['raw_map_points']} raw -> {sum(s['map_points'])} DS points, {s['queries']} queries, {s['iterations']} iterations) |",
                f"| config 4, MAP sharded over {s['world']} GPUs (max over ranks) | {s['map_sharded']['map_voxel_index_ms_max_over_ranks']:.3f} / {s['map_sharded']['registration_device_ms_max_over_ranks']:.3f} / {s['map_sharded']['cycle_ms']:.3f} ms; raw points per rank {s['map_sharded']['raw_points_per_rank']}; pose identical across ranks {s['map_sharded']['pose_bit_identical_across_ranks']}, equal to one GPU {s['map_sharded']['pose_equals_single_gpu']} |",
                f"| config 4, queries sharded, map replicated | registration fused {s['query_sharded']['fused_device_ms_max_over_ranks']:.3f} ms (wall {s['query_sharded']['fused_wall_ms']:.3f}), NCCL-driven {s['query_sharded']['nccl_wall_ms']:.3f} ms; map side {s['query_sharded']['map_voxel_index_ms']:.3f} ms on every rank |",
                f"| cycle speed-up, map sharded vs one GPU | {s['speedup_cycle_vs_single_gpu']:.2f}× |"]
    print("\n".join(out))


if __name__ == "__main__":
    main()
