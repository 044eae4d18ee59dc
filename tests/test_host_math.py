"""The product's device arithmetic header (voxblox_b200/csrc/vbx_math.cuh) compiled for the
host and driven ray by ray (tests/host_sim.cc), against the oracle: every voxel bit-identical.
Also the three-operation division step of the bundle-merge kernel against IEEE division."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import pyoracle as po
from voxblox_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build_sim():
    so = os.path.join(HERE, "libhost_sim.so")
    src = os.path.join(HERE, "host_sim.cc")
    hdr = os.path.join(ROOT, "voxblox_b200", "csrc", "vbx_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               src, "-o", so])
    sim = C.CDLL(so)
    sim.sim_create.restype = C.c_void_p
    sim.sim_count.restype = C.c_uint64
    return sim


def _run(sim, scans, vs, trunc):
    h = C.c_void_p(sim.sim_create())
    om = po.OracleMap(po.OracleLib("port"), po.TsdfConfig(default_truncation_distance=trunc), vs, 16)
    vp = C.c_void_p
    for pts, cols, q, t in scans:
        sim.sim_integrate_simple(h, q.ctypes.data_as(vp), t.ctypes.data_as(vp), pts.ctypes.data_as(vp),
                                 cols.ctypes.data_as(vp), C.c_uint64(len(pts)), C.c_float(vs), C.c_float(trunc),
                                 C.c_float(10000.0), C.c_float(0.1), C.c_float(5.0), 1)
        om.integrate(1, (pts, cols, q, t))
    n = sim.sim_count(h)
    idx = np.zeros((n, 3), np.int32)
    d = np.zeros(n, np.float32)
    w = np.zeros(n, np.float32)
    c = np.zeros(n, np.uint32)
    sim.sim_dump(h, idx.ctypes.data_as(vp), d.ctypes.data_as(vp), w.ctypes.data_as(vp), c.ctypes.data_as(vp))
    sim.sim_destroy(h)
    blocks = om.blocks()
    bad = 0
    for i in range(n):
        b = tuple(int(v) >> 4 for v in idx[i])
        lx, ly, lz = (int(v) & 15 for v in idx[i])
        ov = blocks[b][lx + 16 * (ly + 16 * lz)]
        if ov["distance"] != d[i] or ov["weight"] != w[i] or ov["color"].view(np.uint32)[0] != c[i]:
            bad += 1
    touched = sum(int((v["weight"] > 0).sum()) for v in blocks.values())
    return n, bad, touched


def test_device_math_header_matches_oracle_on_host():
    sim = _build_sim()
    n, bad, touched = _run(sim, [scenes.c1_planar_wall()], 0.2, 0.8)
    assert n >= touched > 2000 and bad == 0
    n, bad, touched = _run(sim, scenes.c3_room_sequence(n_scans=2, width=128, height=96), 0.1, 0.4)
    assert n >= touched > 4000 and bad == 0


def test_three_operation_division_is_correctly_rounded():
    exe = os.path.join(HERE, "exact_div_check")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", os.path.join(HERE, "exact_div_check.c"),
                           "-o", exe, "-lm"])
    out = subprocess.check_output([exe], text=True)
    assert "bad=0" in out, out


def test_merge_form_of_the_ray_walk_equals_the_sequential_walk(tmp_path):
    """k_rays_emit_warp casts a ray as the stable three-way merge of its per-axis crossing chains
    (vbx_math.cuh: dda_rank).  tests/dda_merge_check.cc runs that formulation on the host against
    dda_advance (= RayCaster::nextRayIndex, integrator_utils.cc:106-125) over random, tie-heavy,
    nearly / exactly axis-parallel, clearing and far-from-origin rays: no voxel may differ, and the
    realistic regimes must not need the sequential fallback."""
    exe = str(tmp_path / "dda_merge_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(HERE, "dda_merge_check.cc"),
                           "-o", exe])
    out = subprocess.run([exe, "800000"], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    first = out.stdout.splitlines()[0].split()
    stats = dict(zip(first[0::2], (int(v) for v in first[1::2])))
    assert stats["mismatches"] == 0 and stats["regular"] > 0.7 * stats["rays"]
    per_regime = [int(v) for v in out.stdout.splitlines()[1].split(":")[1].split()]
    assert per_regime[0] == 0 and per_regime[1] == 0 and per_regime[6] == 0   # camera-like and clearing rays


def test_bundle_order_formulation_equals_unordered_map(tmp_path):
    """The Merged integrator's bundle order (voxblox_b200/csrc/vbx_order.cuh) is computed as one
    data-parallel position formula per rehash along libstdc++'s growth schedule.  tests/umap_order_check.cc
    runs that formulation on the host against a real std::unordered_map (the container the reference's
    voxel_map is, tsdf_integrator.cc:318-322).  The device kernels themselves are checked against the same
    container on the GPU (tests/test_order_gpu.py)."""
    exe = str(tmp_path / "umap_order_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(HERE, "umap_order_check.cc"), "-o", exe])
    out = subprocess.check_output([exe]).decode()
    assert " 0 failures" in out, out
