// Device arithmetic of the TSDF hot path: point transform, grid indexing, the
// Amanatides-Woo ray walk and the weighted voxel update.
//
// Results must equal the reference's float32 arithmetic bit for bit, because voxel
// INDEX decisions (floor(x*inv + 1e-6), the DDA's argmin) are made on these values
// (SURVEY.md hard part 2).  Every operation is therefore spelled with the
// round-to-nearest intrinsics (__fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn),
// which the compiler never contracts into FMAs, in the coefficient order of the
// reference's Eigen / minkindr expressions:
//   sum of three coefficients = c0 + (c1 + c2)
//   normalized(v) = v / sqrt(|v|^2)   (v itself when |v|^2 == 0)
//   q * v = v + w*(2 q.vec x v) + q.vec x (2 q.vec x v);   T * p = q * p + t
// The same header compiles for the host (plain IEEE ops; build with
// -ffp-contract=off) so that tests/ can exercise it without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VBX_HD __host__ __device__ __forceinline__
#else
#define VBX_HD inline
#endif

namespace vbx {

VBX_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
VBX_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
VBX_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
VBX_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
VBX_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

struct F3 {
  float x, y, z;
};
VBX_HD F3 f3(float x, float y, float z) {
  F3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
VBX_HD F3 add3(F3 a, F3 b) { return f3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
VBX_HD F3 sub3(F3 a, F3 b) { return f3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
VBX_HD F3 scale3(F3 a, float s) { return f3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
VBX_HD F3 div3(F3 a, float s) { return f3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
VBX_HD float dot3(F3 a, F3 b) {
  return fadd(fmul(a.x, b.x), fadd(fmul(a.y, b.y), fmul(a.z, b.z)));
}
VBX_HD float norm3(F3 a) { return fsqrt(dot3(a, a)); }
VBX_HD F3 unit3(F3 a) {
  const float z = dot3(a, a);
  return z > 0.0f ? div3(a, fsqrt(z)) : a;
}
VBX_HD F3 cross3(F3 a, F3 b) {
  return f3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
            fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}

// T_G_C: unit quaternion (w, x, y, z) and translation.
struct Pose {
  float w, x, y, z;
  F3 t;
};
// kindr QuatTransformation::transform = q.rotate(p) + t, with Eigen's
// QuaternionBase::_transformVector for the rotation (tsdf_integrator.cc:286,356,407).
VBX_HD F3 transform(const Pose& T, F3 p) {
  const F3 qv = f3(T.x, T.y, T.z);
  F3 uv = cross3(qv, p);
  uv = add3(uv, uv);
  return add3(add3(add3(p, scale3(uv, T.w)), cross3(qv, uv)), T.t);
}

struct I3 {
  int x, y, z;
};
VBX_HD I3 i3(int x, int y, int z) {
  I3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}

#define VBX_EPS 1e-6f  // kEpsilon / kFloatEpsilon, core/common.h:139-140

// getGridIndexFromPoint, core/common.h:153-159 / :166-171: floor(x * inv + 1e-6).
// The reference keeps 64-bit indices; the device keeps 32 (|index| < 2^31 voxels).
VBX_HD int grid_coord(float v, float inv) { return (int)floorf(fadd(fmul(v, inv), VBX_EPS)); }
VBX_HD int grid_coord_scaled(float v) { return (int)floorf(fadd(v, VBX_EPS)); }
VBX_HD I3 grid_index(F3 p, float inv) {
  return i3(grid_coord(p.x, inv), grid_coord(p.y, inv), grid_coord(p.z, inv));
}

// getCenterPointFromGridIndex, core/common.h:186-193: the "+ 0.5" literal is a
// double, so the product is formed in double and rounded to float once.
VBX_HD float center_coord(int idx, float grid_size) {
#if defined(__CUDA_ARCH__)
  return __double2float_rn(__dmul_rn(__dadd_rn((double)(float)idx, 0.5), (double)grid_size));
#else
  return (float)(((double)(float)idx + 0.5) * (double)grid_size);
#endif
}

// getVoxelWeight, tsdf_integrator.cc:231-240
VBX_HD float point_weight(float z_C, bool use_const_weight) {
  if (use_const_weight) return 1.0f;
  const float az = fabsf(z_C);
  return az > VBX_EPS ? fdiv(1.0f, fmul(az, az)) : 0.0f;
}

// isPointValid, tsdf_integrator.h:112-129.  Returns 0 invalid, 1 normal, 2 clearing.
// Non-finite points are dropped (the reference front end never passes them,
// voxblox_ros conversions.h:135-137; its RayCaster would read uninitialised state).
VBX_HD int classify_point(F3 p, float min_ray, float max_ray, bool allow_clear, bool freespace) {
  const float r = norm3(p);
  if (!(r == r) || isinf(r)) return 0;
  if (r < min_ray) return 0;
  if (r > max_ray) return (allow_clear || freespace) ? 2 : 0;
  return freespace ? 2 : 1;
}

// ---------------------------------------------------------------------------- colour
// Color::blendTwoColors, core/common.h:105-125 (per channel incl. alpha, C round()).
VBX_HD uint32_t blend_rgba(uint32_t c1, float w1, uint32_t c2, float w2) {
  const float total = fadd(w1, w2);
  w1 = fdiv(w1, total);
  w2 = fdiv(w2, total);
  uint32_t out = 0;
  for (int k = 0; k < 4; ++k) {
    const float a = (float)(int)((c1 >> (8 * k)) & 0xffu);
    const float b = (float)(int)((c2 >> (8 * k)) & 0xffu);
    const float v = fadd(fmul(a, w1), fmul(b, w2));
    const uint32_t q = (uint32_t)(int)roundf(v) & 0xffu;
    out |= q << (8 * k);
  }
  return out;
}

// ------------------------------------------------------------------------ ray caster
// RayCaster, integrator/integrator_utils.cc:72-179, in voxel units.
struct Dda {
  int cx, cy, cz;     // current voxel
  int sx, sy, sz;     // step signs
  float tx, ty, tz;   // t to next boundary
  float dx, dy, dz;   // t step
  unsigned int len;   // emits len + 1 voxels
  unsigned int nx, ny, nz;  // |end voxel - start voxel| per axis (len = nx + ny + nz)
};

VBX_HD int signum_f(float v) { return (v == 0.0f) ? 0 : (v < 0.0f ? -1 : 1); }

VBX_HD void dda_axis(float s, float e, int c, int* sgn, float* tnext, float* tstep) {
  const float r = fsub(e, s);
  const int sg = signum_f(r);
  const float shifted = fsub(s, (float)c);
  const float to_boundary = fsub((float)(sg > 0 ? 1 : 0), shifted);
  *sgn = sg;
  *tnext = fdiv(to_boundary, r);   // +-inf / NaN for an axis-parallel ray, like the reference
  *tstep = fdiv((float)sg, r);
}

// setupRayCaster(start_scaled, end_scaled), integrator_utils.cc:127-179
VBX_HD void dda_setup_scaled(Dda& d, F3 s, F3 e) {
  d.cx = grid_coord_scaled(s.x);
  d.cy = grid_coord_scaled(s.y);
  d.cz = grid_coord_scaled(s.z);
  const int ex = grid_coord_scaled(e.x), ey = grid_coord_scaled(e.y), ez = grid_coord_scaled(e.z);
  d.nx = (unsigned int)abs(ex - d.cx);
  d.ny = (unsigned int)abs(ey - d.cy);
  d.nz = (unsigned int)abs(ez - d.cz);
  d.len = d.nx + d.ny + d.nz;
  dda_axis(s.x, e.x, d.cx, &d.sx, &d.tx, &d.dx);
  dda_axis(s.y, e.y, d.cy, &d.sy, &d.ty, &d.dy);
  dda_axis(s.z, e.z, d.cz, &d.sz, &d.tz, &d.dz);
}

// RayCaster ctor, integrator_utils.cc:72-104
VBX_HD void dda_setup(Dda& d, F3 origin, F3 point_G, bool clearing, bool carving, float max_ray,
                      float voxel_size_inv, float trunc, bool from_origin) {
  const F3 u = unit3(sub3(point_G, origin));
  F3 a, b;
  if (clearing) {
    float l = norm3(sub3(point_G, origin));
    l = fsub(l, trunc);
    l = (l < 0.0f) ? 0.0f : l;        // std::max(l - T, 0)
    l = (max_ray < l) ? max_ray : l;  // std::min(.., max_ray_length_m)
    b = add3(origin, scale3(u, l));
    a = carving ? origin : b;
  } else {
    b = add3(point_G, scale3(u, trunc));
    a = carving ? origin : sub3(point_G, scale3(u, trunc));
  }
  const F3 as = scale3(a, voxel_size_inv), bs = scale3(b, voxel_size_inv);
  if (from_origin) {
    dda_setup_scaled(d, as, bs);
  } else {
    dda_setup_scaled(d, bs, as);
  }
}

// nextRayIndex's advance, integrator_utils.cc:119-122: argmin with the FIRST minimum
// winning (Eigen's minCoeff visitor uses a strict <; NaNs never win).
VBX_HD void dda_advance(Dda& d) {
  int k = 0;
  float m = d.tx;
  if (d.ty < m) {
    m = d.ty;
    k = 1;
  }
  if (d.tz < m) k = 2;
  if (k == 0) {
    d.cx += d.sx;
    d.tx = fadd(d.tx, d.dx);
  } else if (k == 1) {
    d.cy += d.sy;
    d.ty = fadd(d.ty, d.dy);
  } else {
    d.cz += d.sz;
    d.tz = fadd(d.tz, d.dz);
  }
}

// ---- the same walk as a three-way merge (what lets a warp cast one ray cooperatively) ----------
// nextRayIndex picks, len times, the axis whose t_to_next_boundary is smallest (first minimum on
// ties) and adds that axis' t_step to it.  Per axis the visited values form the chain
//   T_a(0) = t_a,  T_a(k+1) = RN(T_a(k) + dt_a)
// which is non-decreasing when dt_a > 0, so the sequence of picks is the stable merge of the
// three chains ordered by (value, axis): element (a, k) is the step with
//   rank = k + #{(b, j): T_b(j) < T_a(k), or T_b(j) == T_a(k) and b < a}
// and after it the walk stands at start + sign_a (k + 1) e_a + sum_b sign_b count_b e_b.  Elements
// can therefore be ranked independently (two binary searches each); only the chains themselves
// are sequential.  Rays with an axis-parallel component (sign 0: the reference's t is -inf / NaN
// there) or non-finite increments are "irregular" and keep the sequential walk.
VBX_HD bool dda_finite_f(float v) { return fabsf(v) <= 3.0e38f; }  // false for inf and NaN
VBX_HD bool dda_is_regular(const Dda& d) {
  return d.sx != 0 && d.sy != 0 && d.sz != 0 && dda_finite_f(d.tx) && dda_finite_f(d.ty) && dda_finite_f(d.tz) &&
         dda_finite_f(d.dx) && dda_finite_f(d.dy) && dda_finite_f(d.dz) && d.dx > 0.0f && d.dy > 0.0f && d.dz > 0.0f;
}
// how many chain elements of an axis are computed: normally the axis takes exactly n steps, so
// elements 0..n-1 are picked and element n is the first one that is not; one more for slack.
// (If that turns out too few, dda_rank reports it and the ray falls back to the sequential walk.)
VBX_HD unsigned int dda_chain_len(unsigned int n_axis, unsigned int len) {
  const unsigned int k = n_axis + 2u;
  return k < len ? k : len;
}
// number of elements of the sorted chain T[0..K) that come before value v; `inclusive`: equal
// values count as before (the other chain belongs to an axis of higher priority)
VBX_HD int dda_count_before(const float* T, int K, float v, bool inclusive) {
  int lo = 0, hi = K;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const bool before = inclusive ? (T[mid] <= v) : (T[mid] < v);
    if (before) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}
// Rank of element (a, k).  counts[b] = steps taken on axis b once this step is done.  Returns
// false if the rank cannot be trusted: some other chain was computed only partially (K_b < len) and
// all of its computed elements precede this one.
VBX_HD bool dda_rank(const float* const T[3], const int K[3], unsigned int len, int a, int k, unsigned int* rank,
                     int counts[3]) {
  const float v = T[a][k];
  bool trusted = true;
  unsigned int r = (unsigned int)k;
  for (int b = 0; b < 3; ++b) {
    if (b == a) {
      counts[b] = k + 1;
      continue;
    }
    const int cb = dda_count_before(T[b], K[b], v, b < a);
    counts[b] = cb;
    r += (unsigned int)cb;
    if (cb == K[b] && (unsigned int)K[b] < len) trusted = false;
  }
  *rank = r;
  return trusted;
}

// ---------------------------------------------------------------------- voxel update
struct TsdfVoxel {  // core/voxel.h:12-16
  float distance;
  float weight;
  uint32_t color;  // r | g << 8 | b << 16 | a << 24 (byte order of struct Color)
};

struct UpdateParams {
  float trunc;        // default_truncation_distance
  float max_weight;
  float voxel_size;
  int use_weight_dropoff;
  int use_sparsity;
  float sparsity_factor;
};

// computeDistance, tsdf_integrator.cc:216-228
VBX_HD float ray_sdf(F3 origin, F3 point_G, int vx, int vy, int vz, float voxel_size) {
  const F3 c = f3(center_coord(vx, voxel_size), center_coord(vy, voxel_size),
                  center_coord(vz, voxel_size));
  const F3 vo = sub3(c, origin), po = sub3(point_G, origin);
  const float dist_G = norm3(po);
  const float dist_G_V = fdiv(dot3(vo, po), dist_G);
  return fsub(dist_G, dist_G_V);
}

// the weight a ray contributes to one voxel: drop-off behind the surface and the
// optional sparsity compensation, tsdf_integrator.cc:161-181
VBX_HD float update_weight(float sdf, float weight, const UpdateParams& P) {
  float w = weight;
  if (P.use_weight_dropoff && sdf < -P.voxel_size) {
    w = fdiv(fmul(weight, fadd(P.trunc, sdf)), fsub(P.trunc, P.voxel_size));
    w = (w < 0.0f) ? 0.0f : w;  // std::max(w, 0.0f)
  }
  if (P.use_sparsity && fabsf(sdf) < P.trunc) w = fmul(w, P.sparsity_factor);
  return w;
}

// the locked read-modify-write of updateTsdfVoxel, tsdf_integrator.cc:186-208
VBX_HD void apply_update(TsdfVoxel& v, float sdf, float w, uint32_t color, const UpdateParams& P) {
  const float new_w = fadd(v.weight, w);
  if (new_w < VBX_EPS) return;
  const float new_sdf = fdiv(fadd(fmul(sdf, w), fmul(v.distance, v.weight)), new_w);
  if (fabsf(sdf) < P.trunc) v.color = blend_rgba(v.color, v.weight, color, w);
  // std::min(T, x) / std::max(-T, x) keep their FIRST argument on NaN
  v.distance = (new_sdf > 0.0f) ? ((new_sdf < P.trunc) ? new_sdf : P.trunc)
                                : ((-P.trunc < new_sdf) ? new_sdf : -P.trunc);
  v.weight = (new_w < P.max_weight) ? new_w : P.max_weight;
}

}  // namespace vbx
