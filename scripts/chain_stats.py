"""Per-voxel update-chain statistics of one bench scan ("why not one thread block per voxel
block"): an approximate numpy ray walk (bundles = voxels of the end points, merged point ~ mean of the bundle,
rays sampled every 5 mm) -- statistics only, not a parity tool."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxblox_b200 import scenes  # noqa: E402

VS, TRUNC = 0.05, 0.2
pts, cols, q, t = scenes.c3_room_scan(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
R = scenes.quat_to_matrix(q.astype(np.float64))
o = t.astype(np.float64)
pg = pts.astype(np.float64) @ R.T + o
r = np.linalg.norm(pts, axis=1)
pg = pg[(r >= 0.1) & (r <= 5.0)]


def pack(v):
    return (v[:, 0] + 4096) + ((v[:, 1] + 4096) << 14) + ((v[:, 2] + 4096) << 28)


v = np.floor(pg / VS + 1e-6).astype(np.int64)
u, inv = np.unique(pack(v), return_inverse=True)
mp = np.zeros((len(u), 3))
np.add.at(mp, inv, pg)
mp /= np.bincount(inv)[:, None]
allv = []
for p in mp:
    d = p - o
    L = np.linalg.norm(d)
    e = o + d / L * (L + TRUNC)
    ts = np.linspace(0, 1, int(np.ceil((L + TRUNC) / 0.005)) + 1)[:, None]
    allv.append(np.unique(pack(np.floor((o + (e - o) * ts) / VS + 1e-6).astype(np.int64))))
allv = np.concatenate(allv)
uu, c = np.unique(allv, return_counts=True)
print(f"rays {len(mp)}  updates ~{len(allv)}  voxels {len(uu)}  mean chain {c.mean():.1f}  p50/p90/p99 "
      f"{np.percentile(c, [50, 90, 99]).tolist()}  longest {c.max()} (the sensor's voxel: one update per ray)")
print(f"voxels with chains > 32: {(c > 32).sum()} holding {100 * c[c > 32].sum() / c.sum():.1f} % of the updates; "
      f"> 256: {(c > 256).sum()} holding {100 * c[c > 256].sum() / c.sum():.1f} %; > 1000: {(c > 1000).sum()}")
bx = ((uu & 0x3fff) - 4096) >> 4
by = (((uu >> 14) & 0x3fff) - 4096) >> 4
bz = ((uu >> 28) - 4096) >> 4
ub, binv = np.unique(bx + 1000 * by + 1000000 * bz, return_inverse=True)
bsum = np.bincount(binv, weights=c)
bmax = np.zeros(len(ub))
np.maximum.at(bmax, binv, c)
order = np.argsort(-bsum)
print(f"blocks {len(ub)}: updates per block (longest chain in it): " +
      ", ".join(f"{int(bsum[i])} ({int(bmax[i])})" for i in order[:5]) + f", ... median {int(np.median(bsum))}")
print(f"share of all updates in the sensor's block: {100 * bsum[order[0]] / bsum.sum():.1f} %; in the top 3 blocks: "
      f"{100 * bsum[order[:3]].sum() / bsum.sum():.1f} %")
