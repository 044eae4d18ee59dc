"""CPU model of ESDF batch propagation policies vs the reference (batch mode: no raise, no seeding)."""
import sys, json, collections
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as po
from voxblox_b200 import scenes
f32 = np.float32

OFF = [(-1,0,0),(1,0,0),(0,-1,0),(0,1,0),(0,0,-1),(0,0,1),(-1,-1,0),(-1,1,0),(1,-1,0),(1,1,0),(0,-1,-1),(0,-1,1),(0,1,-1),(0,1,1),
       (-1,0,-1),(1,0,-1),(-1,0,1),(1,0,1),(-1,-1,-1),(-1,-1,1),(-1,1,-1),(-1,1,1),(1,-1,-1),(1,-1,1),(1,1,-1),(1,1,1)]

def build(voxel, trunc, scans, md, mq):
    lib = po.OracleLib("reference")
    m = po.OracleMap(lib, po.TsdfConfig(default_truncation_distance=trunc, integrator_threads=1), voxel, 16)
    m.esdf_create(po.EsdfConfig(max_distance_m=2.0, default_distance_m=2.0, min_distance_m=trunc / 2, min_diff_m=md, multi_queue=mq))
    for s in scans: m.integrate(2, s)
    m.esdf_update(batch=True)
    return m

class Grid:
    def __init__(self, m, voxel, trunc, md):
        idx = m.block_indices(0)
        self.idx = idx
        lo = idx.min(0); hi = idx.max(0) + 1
        self.lo = lo
        dims = (hi - lo) * 16 + 2          # 1 voxel margin
        self.dims = dims
        self.exists = np.zeros(dims, bool)   # ESDF block exists (batch: every TSDF block)
        self.tsdf_d = np.zeros(dims, f32); self.tsdf_w = np.zeros(dims, f32)
        self.ref_d = np.zeros(dims, f32); self.ref_obs = np.zeros(dims, bool); self.ref_fixed = np.zeros(dims, bool)
        for b in idx:
            o = (b - lo) * 16 + 1
            tv = m.block(b, 0)[0]; ev = m.block(b, 1)[0]
            sl = (slice(o[0], o[0]+16), slice(o[1], o[1]+16), slice(o[2], o[2]+16))
            # linear index x + 16*(y + 16*z) -> array [z][y][x] -> transpose to [x][y][z]
            self.tsdf_d[sl] = tv["distance"].reshape(16,16,16).transpose(2,1,0)
            self.tsdf_w[sl] = tv["weight"].reshape(16,16,16).transpose(2,1,0)
            self.ref_d[sl] = ev["distance"].reshape(16,16,16).transpose(2,1,0)
            self.ref_obs[sl] = ev["observed"].reshape(16,16,16).transpose(2,1,0) != 0
            self.ref_fixed[sl] = ev["fixed"].reshape(16,16,16).transpose(2,1,0) != 0
            self.exists[sl] = True
        self.voxel = f32(voxel); self.min_distance = f32(trunc / 2); self.md = f32(md)
        self.maxd = f32(2.0); self.default = f32(2.0)
        self.dist = [f32(1.0) * self.voxel] * 6 + [f32(np.sqrt(2.0)) * self.voxel] * 12 + [f32(np.sqrt(3.0)) * self.voxel] * 18
        self.dist = [f32(1.0) * self.voxel] * 6 + [f32(f32(np.sqrt(2.0)) * self.voxel)] * 12 + [f32(f32(np.sqrt(3.0)) * self.voxel)] * 8

    def init_batch(self):
        """classification in batch mode on a wiped ESDF: returns d, obs, fixed and the open list in (block sorted, linear) order"""
        obs = self.exists & (self.tsdf_w >= f32(1e-6))
        fixed = obs & (np.abs(self.tsdf_d) < self.min_distance)
        sgn = np.where(self.tsdf_d == 0, f32(0), np.where(self.tsdf_d < 0, f32(-1), f32(1))).astype(f32)
        d = np.where(fixed, self.tsdf_d, sgn * self.default).astype(f32)
        d[~obs] = 0
        open_list = []
        for b in self.idx:
            o = (b - self.lo) * 16 + 1
            f = fixed[o[0]:o[0]+16, o[1]:o[1]+16, o[2]:o[2]+16]
            zz, yy, xx = np.nonzero(f.transpose(2,1,0))   # linear order: x fastest
            for x, y, z in zip(xx, yy, zz): open_list.append((o[0]+x, o[1]+y, o[2]+z))
        return d, obs, fixed, open_list

def signum(v): return f32(0) if v == 0 else (f32(-1) if v < 0 else f32(1))

def relax_pair(G, vd, nd, dist, policy):
    """returns new value for neighbour or None. policy: 'ref' (assignment) or 'dev' (only if nearer)"""
    md = G.md
    if vd > 0 and nd > 0:
        if f32(f32(vd + dist) + md) < nd: return f32(vd + dist)
    elif vd <= 0 and nd <= 0:
        if f32(f32(vd - dist) - md) > nd: return f32(vd - dist)
    else:
        pot = f32(vd - f32(signum(vd) * dist))
        if abs(f32(pot - nd)) > dist:
            nv = pot if signum(pot) == nd else f32(signum(nd) * dist)
            if policy == 'ref': return nv
            if (nv > 0) == (nd > 0) and abs(nv) < abs(nd): return nv
    return None

def run_ref(G, mq, nb=20):
    d, obs, fixed, open_list = G.init_batch()
    inq = np.zeros(G.dims, bool)
    buckets = [collections.deque() for _ in range(nb)]
    state = {"last": 0, "n": 0}
    def push(p, v):
        v = float(v)
        if v > 2.0: v = 2.0
        bi = int(np.floor(abs(v) / 2.0 * (nb - 1)))
        if bi >= nb: bi = nb - 1
        if bi < state["last"]: state["last"] = bi
        buckets[bi].append(p); state["n"] += 1
    for p in open_list:
        inq[p] = True; push(p, d[p])
    pops = 0
    while state["n"]:
        while not buckets[state["last"]]: state["last"] += 1
        p = buckets[state["last"]].popleft(); state["n"] -= 1
        pops += 1
        inq[p] = False
        vd = d[p]
        if not obs[p] or vd >= G.maxd or vd <= -G.maxd: continue
        for i, o in enumerate(OFF):
            q = (p[0]+o[0], p[1]+o[1], p[2]+o[2])
            if not G.exists[q] or not obs[q] or fixed[q]: continue
            nv = relax_pair(G, d[p], d[q], G.dist[i], 'ref')
            if nv is not None:
                d[q] = nv
                if mq or not inq[q]:
                    push(q, nv); inq[q] = True
    return d, obs, pops

def run_dev(G, mq, policy='dev'):
    """level-synchronous sweeps; within a sweep entries processed in list order, reads see earlier writes"""
    d, obs, fixed, front = G.init_batch()
    inq = np.zeros(G.dims, bool)
    for p in front: inq[p] = True
    sweeps = 0
    while front:
        nxt = []
        sweeps += 1
        for p in front:
            inq[p] = False
            vd = d[p]
            if not obs[p] or vd >= G.maxd or vd <= -G.maxd: continue
            for i, o in enumerate(OFF):
                q = (p[0]+o[0], p[1]+o[1], p[2]+o[2])
                if not G.exists[q] or not obs[q] or fixed[q]: continue
                nv = relax_pair(G, vd, d[q], G.dist[i], policy)
                if nv is not None:
                    d[q] = nv
                    if mq or not inq[q]:
                        nxt.append(q); inq[q] = True
        front = nxt
    return d, obs, sweeps

def cmp(G, d, obs, name):
    o = G.ref_obs
    assert (obs == o).all()
    a, b = d[o].astype(np.float64), G.ref_d[o].astype(np.float64)
    err = np.abs(a - b); rel = err / np.maximum(np.abs(b), 1e-3 * float(G.voxel))
    print(name, json.dumps({"observed": int(o.sum()), "bit_exact": round(float((a == b).mean()), 5), "within_1e-4": round(float((rel <= 1e-4).mean()), 5),
                            "max_err": round(float(err.max()), 4), "rmse": round(float(np.sqrt((err**2).mean())), 5)}), flush=True)
    return err


def run_bucket(G, mq, policy, nb=20, shuffle=None):
    """bucket-synchronous: the whole content of the lowest non-empty bucket is one parallel sweep"""
    d, obs, fixed, open_list = G.init_batch()
    inq = np.zeros(G.dims, bool)
    buckets = [[] for _ in range(nb)]
    def bidx(v):
        v = float(v)
        if v > 2.0: v = 2.0
        return min(int(np.floor(abs(v) / 2.0 * (nb - 1))), nb - 1)
    for p in open_list:
        inq[p] = True; buckets[bidx(d[p])].append(p)
    sweeps = 0
    while True:
        b = next((i for i in range(nb) if buckets[i]), None)
        if b is None: break
        front = buckets[b]; buckets[b] = []
        if shuffle is not None: shuffle.shuffle(front)
        sweeps += 1
        for p in front:
            inq[p] = False
            vd = d[p]
            if not obs[p] or vd >= G.maxd or vd <= -G.maxd: continue
            for i, o in enumerate(OFF):
                q = (p[0]+o[0], p[1]+o[1], p[2]+o[2])
                if not G.exists[q] or not obs[q] or fixed[q]: continue
                nv = relax_pair(G, vd, d[q], G.dist[i], policy)
                if nv is not None:
                    d[q] = nv
                    if mq or not inq[q]:
                        buckets[bidx(nv)].append(q); inq[q] = True
    return d, obs, sweeps



def run_jacobi(G, mq, rule, nb=20, bucketed=True, shuffle=None):
    d, obs, fixed, open_list = G.init_batch()
    inq = np.zeros(G.dims, bool)
    buckets = [[] for _ in range(nb)]
    def bidx(v):
        if not bucketed: return 0
        v = float(v)
        if v > 2.0: v = 2.0
        return min(int(np.floor(abs(v) / 2.0 * (nb - 1))), nb - 1)
    for p in open_list:
        inq[p] = True; buckets[bidx(d[p])].append(p)
    sweeps = 0
    while True:
        b = next((i for i in range(nb) if buckets[i]), None)
        if b is None: break
        front = buckets[b]; buckets[b] = []
        if shuffle is not None: shuffle.shuffle(front)
        sweeps += 1
        d0 = d.copy()
        cand_same = {}; cand_mixed = {}
        for p in front:
            inq[p] = False
            vd = d0[p]
            if not obs[p] or vd >= G.maxd or vd <= -G.maxd: continue
            for i, o in enumerate(OFF):
                q = (p[0]+o[0], p[1]+o[1], p[2]+o[2])
                if not G.exists[q] or not obs[q] or fixed[q]: continue
                nd = d0[q]; dist = G.dist[i]
                if (vd > 0 and nd > 0) or (vd <= 0 and nd <= 0):
                    nv = relax_pair(G, vd, nd, dist, 'ref')
                    if nv is not None:
                        if q not in cand_same or abs(nv) < abs(cand_same[q]): cand_same[q] = nv
                else:
                    nv = relax_pair(G, vd, nd, dist, 'ref')
                    if nv is not None: cand_mixed.setdefault(q, []).append(nv)
        for q in set(cand_same) | set(cand_mixed):
            new = d0[q]
            if q in cand_mixed:
                c = cand_mixed[q]
                if rule == 'min': mv = min(c, key=lambda x: abs(x))
                elif rule == 'max': mv = max(c, key=lambda x: abs(x))
                elif rule == 'min_lower_only':
                    mv = min(c, key=lambda x: abs(x))
                    if abs(mv) >= abs(new): mv = new
                new = mv
            if q in cand_same and abs(cand_same[q]) < abs(new): new = cand_same[q]
            if new != d0[q]:
                d[q] = new
                if mq or not inq[q]:
                    buckets[bidx(new)].append(q); inq[q] = True
    return d, obs, sweeps



def replay_dependency_depth(G, mq, nb=20):
    """How parallel could an EXACT replay be?  Two pops commute iff their 3x3x3 footprints are disjoint; the
    longest chain of non-commuting pops (in the reference's pop order) is the number of dependent steps any
    schedule that reproduces the sequential result needs."""
    d, obs, fixed, open_list = G.init_batch()
    inq = np.zeros(G.dims, bool)
    buckets = [collections.deque() for _ in range(nb)]
    state = {"last": 0, "n": 0}

    def push(p, v):
        v = min(float(v), 2.0)
        bi = min(int(np.floor(abs(v) / 2.0 * (nb - 1))), nb - 1)
        state["last"] = min(state["last"], bi)
        buckets[bi].append(p)
        state["n"] += 1
    for p in open_list:
        inq[p] = True
        push(p, d[p])
    last = np.zeros(tuple(np.array(G.dims) + 2), np.int32)
    depth = pops = 0
    while state["n"]:
        while not buckets[state["last"]]:
            state["last"] += 1
        p = buckets[state["last"]].popleft()
        state["n"] -= 1
        pops += 1
        inq[p] = False
        vd = d[p]
        x, y, z = p[0] + 1, p[1] + 1, p[2] + 1
        if not obs[p] or vd >= G.maxd or vd <= -G.maxd:
            last[x, y, z] += 1
            depth = max(depth, int(last[x, y, z]))
            continue
        sl = (slice(x - 1, x + 2), slice(y - 1, y + 2), slice(z - 1, z + 2))
        dj = int(last[sl].max()) + 1
        last[sl] = np.maximum(last[sl], dj)
        depth = max(depth, dj)
        for i, o in enumerate(OFF):
            q = (p[0]+o[0], p[1]+o[1], p[2]+o[2])
            if not G.exists[q] or not obs[q] or fixed[q]:
                continue
            nv = relax_pair(G, d[p], d[q], G.dist[i], 'ref')
            if nv is not None:
                d[q] = nv
                if mq or not inq[q]:
                    push(q, nv)
                    inq[q] = True
    return pops, depth


if __name__ == "__main__":
    import random
    md, mq = (0.0, 1) if len(sys.argv) < 2 or sys.argv[1] == "test" else (1e-3, 0)
    scans = scenes.c3_room_sequence(n_scans=4, width=160, height=120)
    m = build(0.1, 0.4, scans, md, mq)
    G = Grid(m, 0.1, 0.4, md)
    print("room scene, 4 scans of 160x120, voxel 0.1 m, batch update, min_diff", md, "multi_queue", mq)
    d, obs, pops = run_ref(G, mq); cmp(G, d, obs, f"sequential replay of the bucket queue, blocks in sorted order ({pops} pops)")
    d, obs, sw = run_dev(G, mq); cmp(G, d, obs, f"level-synchronous in place, nearest candidate wins, list order ({sw} sweeps)")
    d, obs, sw = run_dev(G, mq, 'ref'); cmp(G, d, obs, f"level-synchronous in place, assignment, list order ({sw} sweeps)")
    for pol in ("ref", "dev"):
        d, obs, sw = run_bucket(G, mq, pol); cmp(G, d, obs, f"bucket by bucket, FIFO inside a bucket, policy {pol} ({sw} sweeps)")
        for seed in (1, 2):
            d, obs, sw = run_bucket(G, mq, pol, shuffle=random.Random(seed)); cmp(G, d, obs, f"bucket by bucket, arbitrary order inside a sweep (seed {seed}), policy {pol}")
    for rule in ("min", "max"):
        d, obs, sw = run_jacobi(G, mq, rule); cmp(G, d, obs, f"bucket by bucket, Jacobi sweeps, conflict rule {rule} ({sw} sweeps)")
    pops, depth = replay_dependency_depth(G, mq)
    print(f"an exact replay: {pops} pops, longest chain of non-commuting pops {depth} (mean {pops / depth:.1f} pops per dependent step)")
