#!/usr/bin/env python
"""Benchmark of the hot path: integrated points/s of MergedTsdfIntegrator on 640x480
scans at 0.05 m voxels (BASELINE.json's metric), one scan per step.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our engine (one JSON line)
  python bench.py --impl reference ...                         the reference's own CPU
                                                               MergedTsdfIntegrator (oracle/_ref)
Multi-GPU runs are launched by torchrun, one rank per GPU (DESIGN.md "multi-GPU").

JSON keys follow the driver's contract; see DESIGN.md "measurement" for how value, e2e,
roofline and cpu_baseline are produced.
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

METRIC = "integrated points/sec (640x480 scan, 0.05 m voxel)"
VOXEL_SIZE = 0.05
TRUNC = 0.2  # 4 voxels (voxblox_ros ros_params.h:66-67, cow_and_lady_dataset.launch)
WORKLOAD = "merged_640x480_room_sequence_0.05m"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_scans(n, start=0):
    from voxblox_b200 import scenes

    return scenes.generate_parallel(scenes.c3_room_scan, range(start, start + n))


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of this workload is
    tens of milliseconds, shorter than one `nvidia-smi -lms` period, so the sampler reads the same
    counters through NVML (nvidia-ml-py) from a thread every ~2 ms; rows are time-stamped and only
    those inside [mark_begin, mark_end] are reported (falling back to nvidia-smi rows if NVML is
    unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index=0):
        self.rows = []      # (t, sm_mhz, max_mhz, reasons bitmask)
        self.proc = None
        self.index = index
        self.stop_flag = False
        self.thread = None
        self.source = None
        self.t_begin = self.t_end = None

    def _nvml_loop(self, nv, h):
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = int(get_reasons(h)) if get_reasons else 0
                self.rows.append((time.perf_counter(), float(sm), float(mx) if mx else None, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                rs = 0
                for nme, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                    if v.lower().startswith("active"):
                        rs |= self.BITS[nme]
                self.rows.append((time.perf_counter(), float(f[0]), float(f[1]), rs))
            except ValueError:
                continue

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        if self.t_end is None:
            self.t_end = time.perf_counter()
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)
        t0 = self.t_begin if self.t_begin is not None else -1e30
        inside = [r for r in self.rows if t0 <= r[0] <= self.t_end]
        window = "timed region"
        if not inside and self.rows:
            # nothing landed inside a very short region: take the sample nearest to it
            mid = 0.5 * (t0 + self.t_end)
            inside = [min(self.rows, key=lambda r: abs(r[0] - mid))]
            window = "nearest sample to the timed region"
        sm = [r[1] for r in inside]
        mx = [r[2] for r in inside if r[2]]
        bits = 0
        for r in inside:
            bits |= r[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(k for k, b in self.BITS.items() if bits & b),
                "source": self.source, "window": window}


# -------------------------------------------------------------------- CPU baseline
def time_reference(scans, threads, kind=2, which=None):
    """Seconds inside integratePointCloud per scan (stopwatch placed like
    voxblox_ros/src/tsdf_server.cc:305-307) for the CPU implementation."""
    from oracle import pyoracle as po

    which = which or ("reference" if po.available("reference") else "port")
    lib = po.OracleLib(which)
    cfg = po.TsdfConfig(default_truncation_distance=TRUNC, integrator_threads=threads)
    m = po.OracleMap(lib, cfg, VOXEL_SIZE, 16)
    secs = []
    for s in scans:
        m.integrate(kind, s)
        secs.append(m.last_seconds())
    m.close()
    return which, secs


def calibrate_threads(scans):
    """The reference defaults to hardware_concurrency threads (tsdf_integrator.h:70) but
    its per-call std::thread fan-out often loses to fewer threads; give it the best."""
    ncpu = os.cpu_count() or 1
    cands = sorted({1, 4, 8, 16, min(32, ncpu), ncpu})
    best = None
    detail = {}
    for t in cands:
        if t > ncpu:
            continue
        _, secs = time_reference(scans, t)
        v = float(np.mean(secs[1:])) if len(secs) > 1 else secs[0]
        detail[str(t)] = round(v * 1e3, 3)
        if best is None or v < best[1]:
            best = (t, v)
    return best[0], detail


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from oracle import pyoracle as po

    scans = make_scans(args.warmup + args.steps)
    threads, detail = calibrate_threads(scans[:3])
    which, secs = time_reference(scans, threads)
    timed = secs[args.warmup:]
    pts = sum(int(s[0].shape[0]) for s in scans[args.warmup:])
    total = float(sum(timed))
    value = pts / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(timed)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "integrator": "merged", "voxel_size_m": VOXEL_SIZE,
                   "truncation_m": TRUNC, "scan": "640x480 pinhole, box room + 4 objects, 15% dropouts",
                   "points_per_scan_mean": pts / max(1, len(timed))},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": which,
                         "sample": f"{len(timed)} scans of the workload; MergedTsdfIntegrator, integrator_threads={threads} "
                                   f"(fastest of {detail} ms/scan; host has {os.cpu_count()} logical cores)"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- ours
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    import voxblox_b200 as vb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total = args.warmup + args.steps
    # weak scaling: every rank integrates its own stream of scans (a different stretch of
    # the trajectory) into its own map -- see DESIGN.md "multi-GPU" for the sharded mode.
    scans = make_scans(n_total, start=rank * n_total)
    npts = [int(s[0].shape[0]) for s in scans]
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=TRUNC)
    opts = vb.EngineOptions(device=local, max_blocks=16384, max_points_per_scan=1 << 19,
                            max_updates_per_pass=1 << 24)

    def fresh():
        layer = vb.Layer(VOXEL_SIZE, 16, engine_options=opts)
        return layer, vb.TsdfIntegratorFactory.create("merged", cfg, layer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) value: inputs resident in HBM, device-timed over the K steps -----------------
    h_xyz = [torch.from_numpy(s[0]).pin_memory() for s in scans]
    h_rgba = [torch.from_numpy(s[1]).pin_memory() for s in scans]
    d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
    d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
    # (1a) the synchronous call: returns when the scan is in the map (one host round trip per scan)
    layer, integ = fresh()
    launches = 0
    for i in range(args.warmup):
        integ.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    barrier()
    layer.timerStart()
    for i in range(args.warmup, n_total):
        integ.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
        launches += integ.counters()["kernel_launches"]
    sync_ms = layer.timerStopMs()
    barrier()
    del layer, integ
    # (1b) headline: the same K scans submitted back to back (vbx_tsdf_integrate_async): the front
    # half of scan i+1 overlaps the back half of scan i; timed until the last scan is in the map
    layer, integ = fresh()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        integ.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    layer.sync()
    barrier()
    launches_before = integ.counters()["kernel_launches_total"]
    sampler.mark_begin()
    layer.timerStart()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_total):
        integ.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    dev_ms = layer.timerStopMs()  # drains both streams first
    sampler.mark_end()
    barrier()
    launches = integ.counters()["kernel_launches_total"] - launches_before
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    last_counters = integ.counters()
    n_blocks = layer.getNumberOfAllocatedBlocks()
    pts_timed = sum(npts[args.warmup:])
    t_ms = torch.tensor([dev_ms, wall_ms, sync_ms], dtype=torch.float64, device=dev)
    tot_pts = torch.tensor([float(pts_timed)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_pts, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, sync_ms = float(t_ms[0]), float(t_ms[1]), float(t_ms[2])
    all_pts = float(tot_pts[0])
    value = all_pts / (dev_ms * 1e-3)

    # ---- (2) e2e: the reference-facing call with HOST buffers (pinned), H2D inside -------
    layer2, integ2 = fresh()
    for rep_ in range(4):  # the GPU idled while the host set up this pass: bring the clocks back up
        for i in range(args.warmup):
            integ.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    hx = [t.numpy() for t in h_xyz]
    hc = [t.numpy() for t in h_rgba]
    for i in range(args.warmup):
        integ2.integratePointCloud((scans[i][2], scans[i][3]), hx[i], hc[i])
    barrier()
    layer2.timerStart()
    for i in range(args.warmup, n_total):
        integ2.integratePointCloud((scans[i][2], scans[i][3]), hx[i], hc[i])
        _ = integ2.counters()  # the step's result block (counters) read back on the host
    e2e_sync_ms = layer2.timerStopMs()
    barrier()
    del layer2, integ2
    # headline e2e: host (page-locked) clouds submitted back to back; every step's H2D copy and the
    # D2H read of its result block (128 B of counters / status) are inside the pipeline
    layer2, integ2 = fresh()
    for i in range(args.warmup):
        integ2.integratePointCloudAsync((scans[i][2], scans[i][3]), hx[i], hc[i])
    layer2.sync()
    barrier()
    layer2.timerStart()
    for i in range(args.warmup, n_total):
        integ2.integratePointCloudAsync((scans[i][2], scans[i][3]), hx[i], hc[i])
    e2e_ms = layer2.timerStopMs()
    barrier()
    # ... and from ordinary pageable memory (what an unmodified caller's Pointcloud is)
    layer2b, integ2b = fresh()
    for i in range(args.warmup):
        integ2b.integratePointCloudAsync((scans[i][2], scans[i][3]), scans[i][0], scans[i][1])
    layer2b.sync()
    barrier()
    layer2b.timerStart()
    for i in range(args.warmup, n_total):
        integ2b.integratePointCloudAsync((scans[i][2], scans[i][3]), scans[i][0], scans[i][1])
    e2e_pageable_ms = layer2b.timerStopMs()
    barrier()
    del layer2b, integ2b
    t_e = torch.tensor([e2e_ms, e2e_sync_ms, e2e_pageable_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = all_pts / (float(t_e[0]) * 1e-3)
    h2d = int(np.mean([16 * n for n in npts[args.warmup:]]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- (3) per-stage device time (profiling pass) and the roofline of the top kernel ---
    layer3, integ3 = fresh()
    U = B = K = 0
    for i in range(n_total):
        if i == args.warmup:
            layer3.setStageProfiling(True)
        integ3.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
        if i >= args.warmup:
            c = integ3.counters()
            U += c["voxels_touched"]
            B += c["blocks_touched"]
            K += c["updates"]
    stages = {k: v for k, v in layer3.stageMs().items() if v[1] > 0}
    top = max(stages, key=lambda k: stages[k][0])
    steps = max(1, args.steps)
    n_mean, u_mean, b_mean, k_mean = pts_timed / steps, U / steps, B / steps, K / steps
    # algorithmic bytes per launch of each stage: SURVEY.md 8(d)'s per-scan
    # figure 16 N + 24 U + 20 B, plus what each stage of THIS design must read and write once.
    alg = {"point_keys": 12 * n_mean + 8 * n_mean, "point_sort": 2 * 8 * n_mean,
           "bundle_merge": (16 + 8) * n_mean, "ray_count": 16 * n_mean, "scan": 8 * n_mean,
           "ray_emit": 8 * k_mean + 20 * b_mean, "assign": 20 * b_mean, "update_sort": 2 * 8 * k_mean,
           "apply": 24 * u_mean + 16 * k_mean}
    # DRAM traffic of the dominant kernels from `ncu --set full` captures (profiles/), per launch
    traffic = {"bundle_merge": 6.35e6, "apply": 4.66e6 + 1.32e6, "ray_emit": 1.18e6}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    top_ms = stages[top][0] / stages[top][1]
    achieved = alg[top] / (top_ms * 1e-3) / 1e9
    scan_alg = 16 * n_mean + 24 * u_mean + 20 * b_mean
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(top),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "alg_bytes_per_launch": alg[top], "avg_launch_ms": top_ms,
                "whole_scan": {"alg_bytes": scan_alg, "ms": dev_ms / steps,
                               "achieved": scan_alg / (dev_ms / steps * 1e-3) / 1e9,
                               "frac": scan_alg / (dev_ms / steps * 1e-3) / 1e9 / peak},
                "stage_ms_per_scan": {k: round(v[0] / steps, 5) for k, v in stages.items()}}

    # ---- (4) CPU baseline on this box's host cores, bounded sample ------------------------
    from oracle import pyoracle as po

    sample = scans[:min(len(scans), 12)]
    threads, detail = calibrate_threads(sample[:3])
    which, secs = time_reference(sample, threads)
    cpu_pts = sum(int(s[0].shape[0]) for s in sample[2:])
    cpu_value = cpu_pts / float(sum(secs[2:]))
    cpu = {"value": cpu_value, "unit": "points/s", "cores": threads, "kind": which,
           "sample": f"{len(sample) - 2} scans of the workload after 2 warm-up scans; MergedTsdfIntegrator "
                     f"with integrator_threads={threads} (fastest of {detail} ms/scan; {os.cpu_count()} logical cores)"}

    # ---- (5) parity spot check against the oracle on the same scans -----------------------
    from tests.parity import compare_tsdf

    layer4, integ4 = fresh()
    omap = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(default_truncation_distance=TRUNC), VOXEL_SIZE, 16)
    for s in scans[:3]:
        integ4.integratePointCloud((s[2], s[3]), s[0], s[1])
        omap.integrate(2, s, order=po.ORDER_CANONICAL)
    rep = compare_tsdf(layer4, omap)

    # ---- (6) the callers either side of the path: ESDF update and
    # incremental mesh after every scan; device time per call, not part of `value`
    downstream = None
    try:
        layer5, integ5 = fresh()
        esdf5 = vb.Layer(VOXEL_SIZE, 16, voxel_type="esdf")
        e5 = vb.EsdfIntegrator(vb.EsdfIntegratorConfig(max_distance_m=2.0, default_distance_m=2.0,
                                                       min_distance_m=TRUNC / 2, min_diff_m=1e-3), layer5, esdf5)
        mesh5 = vb.MeshLayer(layer5.block_size())
        m5 = vb.MeshIntegrator(vb.MeshIntegratorConfig(), layer5, mesh5)
        t_ms5, e_ms5, m_ms5, m_wall5 = [], [], [], []
        k5 = min(n_total, 12)
        for i in range(k5):
            integ5.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
            t_ms5.append(integ5.lastDeviceMs())
            e5.updateFromTsdfLayer(True)
            e_ms5.append(e5.lastDeviceMs())
            t0 = time.perf_counter()
            m5.generateMesh(True, True)
            m_wall5.append((time.perf_counter() - t0) * 1e3)
            m_ms5.append(m5.lastDeviceMs())
        w = slice(min(3, k5 - 1), None)
        downstream = {"scans": k5, "tsdf_sync_device_ms": float(np.mean(t_ms5[w])),
                      "esdf_incremental_device_ms": float(np.mean(e_ms5[w])),
                      "mesh_incremental_device_ms": float(np.mean(m_ms5[w])),
                      "mesh_incremental_wall_ms_incl_download": float(np.mean(m_wall5[w])),
                      "mesh_vertices_last_call": int(m5.last_vertices), "mesh_blocks_last_call": int(m5.last_blocks),
                      "note": "Merged + EsdfIntegrator::updateFromTsdfLayer(true) + MeshIntegrator::generateMesh(true, true) "
                              "after every scan (config C4 + N3); CPU reference: scripts/esdf_bench.py, scripts/mesh_bench.py"}
    except Exception as exc:  # never lose the headline line to an optional section
        downstream = {"error": repr(exc)}

    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "integrator": "merged", "voxel_size_m": VOXEL_SIZE,
                   "truncation_m": TRUNC, "voxels_per_side": 16,
                   "scan": "640x480 pinhole, box room + 4 objects, 15% dropouts, 30 Hz handheld trajectory",
                   "points_per_scan_mean": n_mean, "updates_per_scan_mean": k_mean,
                   "voxels_touched_per_scan_mean": u_mean, "blocks_touched_per_scan_mean": b_mean,
                   "map_blocks_after_run": n_blocks,
                   "l2": "every step integrates a different scan (cloud set > L2); the map's voxel blocks stay "
                         "hot across steps as they do in a mapping session",
                   "parallelism": "one map per GPU" if world > 1 else "single GPU"},
        "clocks": clocks, "wall_ms_per_step": wall_ms / steps,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 128,
                "ms_per_step": float(t_e[0]) / steps, "submission": "vbx_tsdf_integrate_async, page-locked host clouds",
                "synchronous_call": {"value": all_pts / (float(t_e[1]) * 1e-3), "ms_per_step": float(t_e[1]) / steps},
                "pageable_host_memory": {"value": all_pts / (float(t_e[2]) * 1e-3), "ms_per_step": float(t_e[2]) / steps}},
        "submission": "vbx_tsdf_integrate_async: scans queued back to back, timed until the last one is in the map",
        "synchronous_call": {"value": all_pts / (sync_ms * 1e-3), "ms_per_step": sync_ms / steps,
                             "note": "vbx_tsdf_integrate_device, returns when the scan is in the map"},
        "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "parity": {"scans": 3, "blocks_equal": rep["blocks_equal"], "max_rel_err": rep.get("max_rel_err"),
                   "bit_exact_voxels": rep.get("n_bit_exact"), "voxels": rep.get("n_voxels"),
                   "color_mismatch": rep.get("color_mismatch")},
        "last_counters": last_counters,
        "downstream": downstream,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args, rank, world):
    """N > 1.  Headline: the scans are the units and they shard across ranks -- every rank integrates
    its own scan stream into its own map, no collective on the data path (weak scaling).  Also
    measured and reported beside it: the within-scan design of the north_star -- one map, every
    scan sharded over the ranks by contiguous ray ranges with one NCCL all-gather of update records
    per scan (strong scaling; DESIGN.md "multi-GPU")."""
    import torch
    import torch.distributed as dist

    import voxblox_b200 as vb
    from voxblox_b200 import sharded

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_total = args.warmup + args.steps
    scans = make_scans(n_total)  # identical on every rank (seeded)
    npts = [int(s[0].shape[0]) for s in scans]
    cfg = vb.TsdfIntegratorConfig(default_truncation_distance=TRUNC)

    def fresh(rank_, world_):
        opts = vb.EngineOptions(device=local, max_blocks=16384, max_points_per_scan=1 << 19,
                                max_updates_per_pass=1 << 24, rank=rank_, world_size=world_)
        layer = vb.Layer(VOXEL_SIZE, 16, engine_options=opts)
        return layer, vb.TsdfIntegratorFactory.create("merged", cfg, layer)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    d_xyz = [torch.from_numpy(s[0]).to(dev) for s in scans]
    d_rgba = [torch.from_numpy(s[1]).to(dev) for s in scans]
    h_xyz = [torch.from_numpy(s[0]).pin_memory() for s in scans]
    h_rgba = [torch.from_numpy(s[1]).pin_memory() for s in scans]

    # ---- value: inputs resident in HBM on every rank
    layer, integ = fresh(rank, world)
    sh = sharded.ShardedTsdfIntegrator(integ, record_capacity=(1 << 24) // world)
    for i in range(args.warmup):
        sh.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    sampler.mark_begin()
    layer.timerStart()
    launches = 0
    xbytes = 0
    for i in range(args.warmup, n_total):
        sh.integratePointCloudDevice((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
        launches += integ.counters()["kernel_launches"]
        xbytes += sh.last_exchange_bytes
    dev_ms = layer.timerStopMs()
    sampler.mark_end()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms[0])
    pts_timed = float(sum(npts[args.warmup:]))
    value = pts_timed / (dev_ms * 1e-3)

    # ---- e2e: host buffers, H2D of the cloud on every rank inside the timed region
    layer2, integ2 = fresh(rank, world)
    sh2 = sharded.ShardedTsdfIntegrator(integ2, record_capacity=(1 << 24) // world)
    stage_xyz = torch.empty((max(npts), 3), dtype=torch.float32, device=dev)
    stage_rgba = torch.empty((max(npts), 4), dtype=torch.uint8, device=dev)

    def e2e_step(i):
        stage_xyz[:npts[i]].copy_(h_xyz[i], non_blocking=True)
        stage_rgba[:npts[i]].copy_(h_rgba[i], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        sh2.integratePointCloudDevice((scans[i][2], scans[i][3]), stage_xyz.data_ptr(), stage_rgba.data_ptr(), npts[i])
        return integ2.counters()

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_total):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t_e = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = pts_timed / (float(t_e[0]) * 1e-3)

    # ---- replicas (weak scaling, for context): every rank integrates its own scan stream
    layer3, integ3 = fresh(0, 1)
    sampler3 = ClockSampler(local)
    if rank == 0:
        sampler3.start()
    for i in range(args.warmup):
        integ3.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    layer3.sync()
    barrier()
    l3 = integ3.counters()["kernel_launches_total"]
    sampler3.mark_begin()
    layer3.timerStart()
    for i in range(args.warmup, n_total):
        integ3.integratePointCloudAsync((scans[i][2], scans[i][3]), d_xyz[i].data_ptr(), d_rgba[i].data_ptr(), npts[i])
    rep_ms = layer3.timerStopMs()
    sampler3.mark_end()
    barrier()
    rep_launches = integ3.counters()["kernel_launches_total"] - l3
    clocks = sampler3.stop() if rank == 0 else None
    t_r = torch.tensor([rep_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
    replicas_value = world * pts_timed / (float(t_r[0]) * 1e-3)

    # all replicas must hold the same map: compare a digest of rank 0's blocks with every rank's
    import hashlib

    idx = layer.getAllAllocatedBlocks()
    vox, _ = layer.getBlocks(idx)
    digest = int.from_bytes(hashlib.sha256(idx.tobytes() + vox.tobytes()).digest()[:7], "little")
    dg = torch.tensor([digest], dtype=torch.int64, device=dev)
    allg = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allg, dg)
    replicas_identical = bool((allg == allg[0]).all())

    # ---- replicas, e2e: every rank feeds its own map from pinned host buffers
    layer4, integ4 = fresh(0, 1)
    hx = [t.numpy() for t in h_xyz]
    hc = [t.numpy() for t in h_rgba]
    for i in range(args.warmup):
        integ4.integratePointCloudAsync((scans[i][2], scans[i][3]), hx[i], hc[i])
    layer4.sync()
    barrier()
    layer4.timerStart()
    for i in range(args.warmup, n_total):
        integ4.integratePointCloudAsync((scans[i][2], scans[i][3]), hx[i], hc[i])
    rep_e2e_ms = layer4.timerStopMs()
    barrier()
    t_re = torch.tensor([rep_e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t_re, op=dist.ReduceOp.MAX)
    replicas_e2e = world * pts_timed / (float(t_re[0]) * 1e-3)
    n_l = torch.tensor([rep_launches], dtype=torch.int64, device=dev)
    dist.all_reduce(n_l, op=dist.ReduceOp.SUM)

    if rank == 0:
        steps = max(1, args.steps)
        line = {
            "metric": METRIC, "value": replicas_value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(t_r[0]) / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "integrator": "merged", "voxel_size_m": VOXEL_SIZE,
                       "truncation_m": TRUNC, "voxels_per_side": 16,
                       "scan": "640x480 pinhole, box room + 4 objects, 15% dropouts, 30 Hz handheld trajectory",
                       "points_per_scan_mean": pts_timed / steps,
                       "parallelism": f"{world} replicas: one map and one scan stream per GPU, scans sharded across "
                                      "ranks, no collective on the data path (a 0.6 ms scan is too small to shard "
                                      "internally; see ray_range_sharded)",
                       "l2": "every step integrates a different scan; the map's blocks stay hot"},
            "clocks": clocks,
            "e2e": {"value": replicas_e2e, "unit": "points/s", "h2d_bytes_per_step": int(16 * pts_timed / steps),
                    "d2h_bytes_per_step": 128, "ms_per_step": float(t_re[0]) / steps,
                    "submission": "vbx_tsdf_integrate_async, page-locked host clouds"},
            "submission": "vbx_tsdf_integrate_async: scans queued back to back, timed until the last one is in the map",
            "gpu_launches": int(n_l[0]),
            # the north_star's within-scan design: ONE map, every scan sharded by contiguous ray ranges,
            # one NCCL all-gather of update records per scan; replicas bit-identical.  Strong scaling.
            "ray_range_sharded": {"value": value, "unit": "points/s", "scaling": "strong",
                                  "ms_per_step": dev_ms / steps,
                                  "e2e": {"value": e2e_value, "ms_per_step": float(t_e[0]) / steps},
                                  "exchange_bytes_per_scan": xbytes / steps, "gpu_launches": int(launches),
                                  "replicas_identical": replicas_identical},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import __graft_entry__ as g

    if rank == 0:
        with open(os.devnull, "w") as devnull:
            old = sys.stdout
            sys.stdout = sys.stderr
            try:
                g.build()
            finally:
                sys.stdout = old
    if args.impl == "reference":
        run_reference_arm(args, rank)
    elif world > 1:
        run_sharded(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
