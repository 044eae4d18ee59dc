"""Generates tests/golden/digests.json: SHA-256 digests of the layers the REFERENCE's own
integrator sources (oracle/_ref/libvbx_ref.so, built from /root/reference by oracle/Makefile)
produce on small seeded scans.  Run it where /root/reference exists:

    python tests/golden/make_golden.py

The digests pin the CPU restatement (oracle/vbx_oracle.cc) on machines without the reference
(tests/test_oracle_pin.py::test_port_matches_committed_golden).  The reference has no golden
files of its own for TSDF / ESDF values, so these are the vectors.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from voxblox_b200 import scenes  # noqa: E402

CASES = {
    # name: (kind, scans, voxel_size, trunc, tsdf kwargs, esdf kwargs or None, batch)
    "c1_simple": (1, "c1", 0.2, 0.8, {}, None, False),
    "c1_merged": (2, "c1", 0.2, 0.8, {}, None, False),
    "c1_fast": (3, "c1", 0.2, 0.8, {}, None, False),
    "room_simple": (1, "room", 0.1, 0.4, {}, None, False),
    "room_merged_esdf_incremental": (2, "room", 0.1, 0.4, {},
                                     dict(max_distance_m=4.0, default_distance_m=4.0, min_distance_m=0.2,
                                          min_diff_m=0.0, multi_queue=1), False),
    "room_merged_esdf_batch_defaults": (2, "room", 0.1, 0.4, {}, dict(min_distance_m=0.2), True),
    "room_fast_sorted": (3, "room", 0.1, 0.4, dict(integration_order_mode=1), None, False),
    "room_merged_antigrazing_constweight": (2, "room", 0.1, 0.4, dict(enable_anti_grazing=1, use_const_weight=1),
                                            None, False),
    "room_simple_nocarving_shortrange": (1, "room", 0.1, 0.4, dict(voxel_carving_enabled=0, max_ray_length_m=2.0),
                                         None, False),
    # test_clear_spheres.cc:107-165: addNewRobotPosition before every scan, incremental updates
    "room_merged_esdf_clear_spheres": (2, "room", 0.1, 0.4, {},
                                       dict(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=0.2,
                                            min_diff_m=0.0, clear_sphere_radius=1.0, occupied_sphere_radius=2.5),
                                       False, "robot"),
}


def case_scans(name):
    if name == "c1":
        return [scenes.c1_planar_wall()]
    return scenes.c3_room_sequence(n_scans=3, width=96, height=72)


def layer_digest(omap, layer):
    h = hashlib.sha256()
    idx = omap.block_indices(layer)
    h.update(idx.tobytes())
    for i in idx:
        vox, upd = omap.block(i, layer)
        h.update(vox.tobytes())
        h.update(bytes([upd]))
    return h.hexdigest(), int(len(idx))


def run_case(lib, name):
    kind, scan_name, vs, trunc, tkw, ekw, batch = CASES[name][:7]
    robot = len(CASES[name]) > 7
    omap = po.OracleMap(lib, po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1, **tkw), vs, 16)
    if ekw is not None:
        omap.esdf_create(po.EsdfConfig(**ekw))
    for s in case_scans(scan_name):
        if robot:
            omap.esdf_add_robot_position(s[3])
        omap.integrate(kind, s)
        if ekw is not None and not batch:
            omap.esdf_update(batch=False, clear_updated_flag=True)
    if ekw is not None and batch:
        omap.esdf_update(batch=True)
    out = {"tsdf": layer_digest(omap, 0)}
    if ekw is not None:
        out["esdf"] = layer_digest(omap, 1)
    return out


def scans_digest(scan_name):
    h = hashlib.sha256()
    for s in case_scans(scan_name):
        for a in s:
            h.update(a.tobytes())
    return h.hexdigest()


def main():
    lib = po.OracleLib("reference")
    digests = {name: run_case(lib, name) for name in CASES}
    inputs = {n: scans_digest(n) for n in ("c1", "room")}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "digests.json")
    with open(path, "w") as f:
        json.dump({"generated_by": "oracle/_ref/libvbx_ref.so (reference sources @ /root/reference)",
                   "inputs": inputs, "digests": digests}, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
