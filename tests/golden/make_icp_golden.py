"""Generates tests/golden/icp.json: refined poses and fused-batch counts of the REFERENCE's own
voxblox::ICP (src/alignment/icp.cc compiled where it lies into oracle/_ref/libvbx_ref.so, one thread) on
small seeded inputs.  Run it where /root/reference exists:

    python tests/golden/make_icp_golden.py

The values pin the restatement's ICP (oracle/vbx_oracle.cc) on machines without the reference
(tests/test_icp_cpu.py).  The reference has no test or golden vector of its own for ICP (SURVEY.md
section 4), so these are the vectors.  Floats are stored as exact hex strings."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from tests.golden import make_golden as mg  # noqa: E402

CASES = {
    # name: (icp kwargs, pose perturbation (dx, dy, dz), seed)
    "yaw_only_true_pose": (dict(), (0.0, 0.0, 0.0), 1),
    "yaw_only_perturbed": (dict(), (0.06, -0.05, 0.03), 2),
    "full_rotation_perturbed": (dict(refine_roll_pitch=1), (0.06, -0.05, 0.03), 3),
    "big_batches_low_ratio": (dict(mini_batch_size=50, min_match_ratio=0.5, subsample_keep_ratio=0.9), (-0.04, 0.02, 0.0), 4),
}


def run_case(lib, name):
    kw, dt, seed = CASES[name]
    scans = mg.case_scans("room")
    omap = po.OracleMap(lib, po.TsdfConfig(default_truncation_distance=0.4, integrator_threads=1), 0.1, 16)
    for s in scans[:2]:
        omap.integrate(po.MERGED, s)
    s = scans[2]
    t0 = (np.asarray(s[3], np.float64) + np.asarray(dt)).astype(np.float32)
    q, t, n = omap.icp(po.IcpConfig(**kw), s[0], s[2], t0, seed)
    return q, t, n


def main():
    lib = po.OracleLib("reference")
    out = {}
    for name in CASES:
        q, t, n = run_case(lib, name)
        out[name] = {"q_wxyz": [float(v).hex() for v in q], "t": [float(v).hex() for v in t], "num_updates": n}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp.json")
    with open(path, "w") as f:
        json.dump({"generated_by": "oracle/_ref/libvbx_ref.so (reference icp.cc @ /root/reference, num_threads = 1)",
                   "inputs": {"room": mg.scans_digest("room")}, "cases": out}, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
