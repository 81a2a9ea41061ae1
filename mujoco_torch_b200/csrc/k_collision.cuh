// k_collision: one thread per (candidate pair, env); grid.x walks the static pair table, grid.y the env blocks.
// Narrowphase for plane/sphere/capsule primitives (reference collision_primitive.py:29-75,172-221 and the
// geometry helpers math.py:496-569).  Each pair writes its contacts straight to their final slots: the
// reference's concat + argsort-by-condim (collision_driver.py:815-850) is a static permutation baked into
// pair_slot at device_put time, so nothing is sorted on the device.
#pragma once
#include "mjt_common.cuh"
#include "k_smooth.cuh"

template <typename T>
struct GeomPose {
  V3<T> pos;
  T mat[9];
};

template <typename T>
__device__ __forceinline__ void load_geom(const ModelView<T>& m, const EnvBuf<T>& d, int g, GeomPose<T>& o) {
  ld3(d, m.D(D_geom_xpos) + 3 * g, o.pos);
#pragma unroll
  for (int k = 0; k < 9; ++k) o.mat[k] = d.get(m.D(D_geom_xmat) + 9 * g + k);
}
template <typename T> __device__ __forceinline__ V3<T> zaxis(const GeomPose<T>& g) { return V3<T>(g.mat[2], g.mat[5], g.mat[8]); }

template <typename T>
__device__ __forceinline__ void write_contact(const ModelView<T>& m, const EnvBuf<T>& d, int slot, T dist, const V3<T>& pos, const T* frame) {
  if (slot < 0) return;
  // max_contact_points models: `slot` is the candidate's index and the contact goes to the candidate scratch (k_topk selects)
  const bool tk = m.h(H_topk) > 0;
  const int oD = m.D(tk ? D_scr_cand_dist : D_contact_dist), oP = m.D(tk ? D_scr_cand_pos : D_contact_pos);
  const int oF = m.D(tk ? D_scr_cand_frame : D_contact_frame);
  d.set(oD + slot, dist);
  st3(d, oP + 3 * slot, pos);
#pragma unroll
  for (int k = 0; k < 9; ++k) d.set(oF + 9 * slot + k, frame[k]);
}

// reference _plane_sphere :29-38
template <typename T>
__device__ __forceinline__ void plane_sphere_core(const V3<T>& n, const V3<T>& ppos, const V3<T>& spos, T r, T* dist, V3<T>* pos) {
  *dist = dot(spos - ppos, n) - r;
  *pos = spos - n * (r + T(0.5) * (*dist));
}
// reference _sphere_sphere :172-186
template <typename T>
__device__ __forceinline__ void sphere_sphere_core(const V3<T>& p1, T r1, const V3<T>& p2, T r2, T* dist, V3<T>* pos, V3<T>* n) {
  T len;
  V3<T> nn = normalize3(p2 - p1, &len);
  if (len == T(0)) nn = V3<T>(1, 0, 0);
  *dist = len - (r1 + r2);
  *pos = p1 + nn * (r1 + (*dist) * T(0.5));
  *n = nn;
}
// reference math.closest_segment_point :506-510
template <typename T>
__device__ __forceinline__ V3<T> closest_segment_point(const V3<T>& a, const V3<T>& b, const V3<T>& pt) {
  const V3<T> ab = b - a;
  T t = dot(pt - a, ab) / (dot(ab, ab) + T(1e-6));
  t = t_clamp(t, T(0), T(1));
  return a + ab * t;
}
// reference math.closest_segment_to_segment_points :523-569
template <typename T>
__device__ __forceinline__ void closest_segment_to_segment(const V3<T>& a0, const V3<T>& a1, const V3<T>& b0, const V3<T>& b1, V3<T>* oa, V3<T>* ob) {
  T len_a, len_b;
  const V3<T> dir_a = normalize3(a1 - a0, &len_a), dir_b = normalize3(b1 - b0, &len_b);
  const T ha = len_a * T(0.5), hb = len_b * T(0.5);
  const V3<T> a_mid = a0 + dir_a * ha, b_mid = b0 + dir_b * hb;
  const V3<T> trans = a_mid - b_mid;
  const T dadb = dot(dir_a, dir_b), dat = dot(dir_a, trans), dbt = dot(dir_b, trans);
  const T denom = T(1) - dadb * dadb;
  const T ta0 = (-dat + dadb * dbt) / (denom + T(1e-6));
  const T tb0 = dbt + ta0 * dadb;
  const T ta = t_max(t_min(ta0, ha), -ha), tb = t_max(t_min(tb0, hb), -hb);
  V3<T> best_a = a_mid + dir_a * ta, best_b = b_mid + dir_b * tb;
  const V3<T> new_a = closest_segment_point(a0, a1, best_b), new_b = closest_segment_point(b0, b1, best_a);
  const V3<T> da = best_b - new_a, db = best_a - new_b;
  const T d1 = dot(da, da), d2 = dot(db, db);
  if (d1 < d2) best_a = new_a; else best_b = new_b;
  *oa = best_a;
  *ob = best_b;
}

template <typename T>
__device__ void convex_pair(const ModelView<T>& m, const EnvBuf<T>& d, int fn, int g1, int g2, const int* slot);

// height fields (k_hfield.cuh): a real call, so that its stack and registers do not weigh on the other pair types
template <typename T>
__device__ __noinline__ void hfield_pair(const ModelView<T>& m, const EnvBuf<T>& d, int p, int fn, int g1, int g2, const int* slot);

// HF: the instantiation for models with height-field pairs (its callee costs registers and stack; everything else runs HF = false)
template <typename T, bool HF>
__device__ void collide_pair(const ModelView<T>& m, const EnvBuf<T>& d, int p) {
  const int fn = m.I(I_pair_fn)[p], g1 = m.I(I_pair_g1)[p], g2 = m.I(I_pair_g2)[p];
  const int* slot = m.I(I_pair_slot) + 4 * p;
  const T* size = m.F(F_geom_size);
  if (fn >= FN_HFIELD_SPHERE) {
    if (HF) hfield_pair<T>(m, d, p, fn, g1, g2, slot);
    return;
  }
  if (fn >= FN_PLANE_CONVEX) {
    convex_pair<T>(m, d, fn, g1, g2, slot);
    return;
  }
  GeomPose<T> a, b;
  load_geom(m, d, g1, a);
  load_geom(m, d, g2, b);
  T frame[9], dist;
  V3<T> pos, n;
  switch (fn) {
    case FN_PLANE_SPHERE: {
      n = zaxis(a);
      plane_sphere_core(n, a.pos, b.pos, size[3 * g2], &dist, &pos);
      make_frame(n, frame);
      write_contact(m, d, slot[0], dist, pos, frame);
    } break;
    case FN_PLANE_CAPSULE: {  // :48-74
      n = zaxis(a);
      const V3<T> axis = zaxis(b);
      T bn;
      V3<T> bb = normalize3(axis - n * dot(n, axis), &bn);
      if (bn < T(0.5)) bb = (n.y > T(-0.5) && n.y < T(0.5)) ? V3<T>(0, 1, 0) : V3<T>(0, 0, 1);
      const V3<T> c = cross(n, bb);
      frame[0] = n.x; frame[1] = n.y; frame[2] = n.z; frame[3] = bb.x; frame[4] = bb.y; frame[5] = bb.z;
      frame[6] = c.x; frame[7] = c.y; frame[8] = c.z;
      const V3<T> seg = axis * size[3 * g2 + 1];
      plane_sphere_core(n, a.pos, b.pos + seg, size[3 * g2], &dist, &pos);
      write_contact(m, d, slot[0], dist, pos, frame);
      plane_sphere_core(n, a.pos, b.pos - seg, size[3 * g2], &dist, &pos);
      write_contact(m, d, slot[1], dist, pos, frame);
    } break;
    case FN_SPHERE_SPHERE: {
      sphere_sphere_core(a.pos, size[3 * g1], b.pos, size[3 * g2], &dist, &pos, &n);
      make_frame(n, frame);
      write_contact(m, d, slot[0], dist, pos, frame);
    } break;
    case FN_SPHERE_CAPSULE: {  // :195-201
      const V3<T> seg = zaxis(b) * size[3 * g2 + 1];
      const V3<T> pt = closest_segment_point(b.pos - seg, b.pos + seg, a.pos);
      sphere_sphere_core(a.pos, size[3 * g1], pt, size[3 * g2], &dist, &pos, &n);
      make_frame(n, frame);
      write_contact(m, d, slot[0], dist, pos, frame);
    } break;
    case FN_CAPSULE_CAPSULE: {  // :204-221
      const V3<T> s1 = zaxis(a) * size[3 * g1 + 1], s2 = zaxis(b) * size[3 * g2 + 1];
      V3<T> p1, p2;
      closest_segment_to_segment(a.pos - s1, a.pos + s1, b.pos - s2, b.pos + s2, &p1, &p2);
      sphere_sphere_core(p1, size[3 * g1], p2, size[3 * g2], &dist, &pos, &n);
      make_frame(n, frame);
      write_contact(m, d, slot[0], dist, pos, frame);
    } break;
    default: break;
  }
}

template <typename T, bool HF>
__global__ void __launch_bounds__(MJT_ENV_THREADS, MJT_COLLISION_MINB) k_collision(const int* __restrict__ gblob, int blob_words, T* __restrict__ out, int B) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  // grid.x walks the candidate pairs (fastest), grid.y the env blocks (see k_constraint)
  for (int env = blockIdx.y * blockDim.x + threadIdx.x; env < B; env += gridDim.y * blockDim.x) {
    EnvBuf<T> d{out + env, (size_t)B};
    for (int p = blockIdx.x; p < m.h(H_npair_cand); p += gridDim.x) collide_pair<T, HF>(m, d, p);
  }
}


// k_topk: max_contact_points (reference collision_driver.collision :822-840).  One thread per env: of the ncand candidate contacts
// the K with the smallest distance are kept, in ascending distance like torch.topk(-dist, K) returns them (ties: the lower
// candidate index first), then permuted by the baked argsort of the survivors' (equal) condims.  The contact and the index of the
// candidate it came from go to the final slots; k_constraint looks the contact's parameters and bodies up through that index.
template <typename T>
__global__ void __launch_bounds__(MJT_ENV_THREADS) k_topk(const int* __restrict__ gblob, int blob_words, T* __restrict__ out, int B) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  EnvBuf<T> d{out + env, (size_t)B};
  const int K = m.h(H_topk), n = m.h(H_ncand);
  const int oC = m.D(D_scr_cand_dist), oCP = m.D(D_scr_cand_pos), oCF = m.D(D_scr_cand_frame), oS = m.D(D_scr_consrc);
  const int* perm = m.I(I_topk_perm);
  T last = -t_inf<T>();
  int last_i = -1;
  for (int k = 0; k < K; ++k) {
    // next candidate in (dist, index) order after (last, last_i)
    T best = t_inf<T>();
    int bi = -1;
    for (int c = 0; c < n; ++c) {
      const T v = d.get(oC + c);
      const bool after = (v > last) || (v == last && c > last_i);
      if (after && (bi < 0 || v < best)) { best = v; bi = c; }
    }
    if (bi < 0) { bi = last_i < 0 ? 0 : last_i; best = d.get(oC + bi); }   // NaN distances: keep the output defined
    last = best;
    last_i = bi;
    // position of this survivor after the condim sort: slot s with perm[s] == k
    int slot = k;
    for (int s = 0; s < K; ++s) if (perm[s] == k) slot = s;
    d.set(m.D(D_contact_dist) + slot, best);
    for (int j = 0; j < 3; ++j) d.set(m.D(D_contact_pos) + 3 * slot + j, d.get(oCP + 3 * bi + j));
    for (int j = 0; j < 9; ++j) d.set(m.D(D_contact_frame) + 9 * slot + j, d.get(oCF + 9 * bi + j));
    d.set(oS + slot, T(bi));
  }
}
