// k_convex: convex (box / mesh) narrowphase -- plane_convex, sphere_convex, capsule_convex, convex_convex.
// One thread per (pair, env), called from k_collision.  Follows reference collision_convex.py: _manifold_points :183-235,
// _clip_edge_to_planes :265-327, _clip :330-392, _create_contact_manifold :395-449, _sat_hull_hull :464-601,
// plane_convex :604-623, sphere_convex :626-699, capsule_convex :702-802, convex_convex :805-856.  The reference's
// argmax/argmin "first index wins" rule is kept everywhere (strict comparisons while scanning in index order).
// Hull tables (vertices, padded CCW polygon faces, face normals, shared edges) are read from the staged ModelBlob.
#pragma once
#include "k_collision.cuh"

#define MJT_MAXFV 20  // reference mesh.py:32 _MAX_HULL_FACE_VERTICES
#define MJT_VCACHE 32 // transformed hull-A vertices kept in local memory; larger hulls re-transform on the fly

template <typename T>
struct Hull {
  const T* vert;
  const T* norm;
  const int* face;
  const int* edge;
  int nvert, nface, nfv, nedge;
  __device__ __forceinline__ V3<T> v(int i) const { return V3<T>(vert[3 * i], vert[3 * i + 1], vert[3 * i + 2]); }
  __device__ __forceinline__ V3<T> n(int f) const { return V3<T>(norm[3 * f], norm[3 * f + 1], norm[3 * f + 2]); }
  // k-th corner of face f, padded by repeating the last corner (collision_convex.py:815-823)
  __device__ __forceinline__ int fv(int f, int k) const { return face[f * nfv + (k < nfv ? k : nfv - 1)]; }
};

template <typename T>
__device__ __forceinline__ Hull<T> load_hull(const ModelView<T>& m, int g) {
  const int c = m.I(I_geom_cvx)[g];
  Hull<T> h;
  h.vert = m.F(F_cvx_vert) + 3 * m.I(I_cvx_vertadr)[c];
  h.norm = m.F(F_cvx_norm) + 3 * m.I(I_cvx_normadr)[c];
  h.face = m.I(I_cvx_face) + m.I(I_cvx_faceadr)[c];
  h.edge = m.I(I_cvx_edge) + 2 * m.I(I_cvx_edgeadr)[c];
  h.nvert = m.I(I_cvx_nvert)[c];
  h.nface = m.I(I_cvx_nface)[c];
  h.nfv = m.I(I_cvx_nfv)[c];
  h.nedge = m.I(I_cvx_nedge)[c];
  return h;
}

template <typename T> __device__ __forceinline__ V3<T> mat_mul(const T* r, const V3<T>& v) {
  return V3<T>(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z);
}
template <typename T> __device__ __forceinline__ V3<T> mat_tmul(const T* r, const V3<T>& v) {
  return V3<T>(r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z);
}
template <typename T> __device__ __forceinline__ V3<T> ldv(const T* a, int i) { return V3<T>(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
template <typename T> __device__ __forceinline__ void stv(T* a, int i, const V3<T>& v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

// _closest_segment_point_plane :39-63
template <typename T>
__device__ __forceinline__ V3<T> seg_plane(const V3<T>& a, const V3<T>& b, const V3<T>& p0, const V3<T>& n) {
  const T dd = dot(p0, n);
  const V3<T> ab = b - a;
  const T denom = dot(n, ab);
  T t = (dd - dot(n, a)) / (denom + (denom == T(0) ? T(1e-6) : T(0)));
  t = t_clamp(t, T(0), T(1));
  return a + ab * t;
}

// _manifold_points :183-235 over N points pt(i) with validity mk(i)
template <typename T, typename P, typename K>
__device__ __forceinline__ void manifold_points(int N, P pt, K mk, const V3<T>& norm, int* idx) {
  const T NEG = T(-1e6);
  int ai = 0;
  {
    T best = mk(0) ? T(0) : NEG;
    for (int i = 1; i < N; ++i) { const T v = mk(i) ? T(0) : NEG; if (v > best) { best = v; ai = i; } }
  }
  const V3<T> a = pt(ai);
  int bi = 0;
  {
    T best = -t_inf<T>();
    for (int i = 0; i < N; ++i) {
      const V3<T> q = a - pt(i);
      const T v = dot(q, q) + (mk(i) ? T(0) : NEG);
      if (i == 0 || v > best) { best = v; bi = i; }
    }
  }
  const V3<T> b = pt(bi);
  const V3<T> ab = cross(norm, a - b);
  int ci = 0;
  {
    T best = T(0);
    for (int i = 0; i < N; ++i) {
      const T v = t_abs(dot(a - pt(i), ab)) + (mk(i) ? T(0) : NEG);
      if (i == 0 || v > best) { best = v; ci = i; }
    }
  }
  const V3<T> c = pt(ci);
  const V3<T> ac = cross(norm, a - c), bc = cross(norm, b - c);
  int di = 0;
  {
    T best = T(0);
    for (int i = 0; i < N; ++i) {
      const T v = t_abs(dot(b - pt(i), bc)) + (mk(i) ? T(0) : NEG);
      if (i == 0 || v > best) { best = v; di = i; }
    }
    for (int i = 0; i < N; ++i) {
      const T v = t_abs(dot(a - pt(i), ac)) + (mk(i) ? T(0) : NEG);
      if (v > best) { best = v; di = i; }
    }
  }
  idx[0] = ai; idx[1] = bi; idx[2] = ci; idx[3] = di;
}

// _clip_edge_to_planes :265-327; planes k = 0..P-1 given by plane(k, &point, &normal).  Returns the mask.
template <typename T, typename PL>
__device__ __forceinline__ bool clip_edge_to_planes(const V3<T>& p0, const V3<T>& p1, int P, PL plane, V3<T>* o0, V3<T>* o1) {
  const V3<T> d01 = p1 - p0, d10 = p0 - p1;
  V3<T> n0 = p0, n1 = p1;
  T b0 = T(0), b1 = T(0);
  bool both = false;
  for (int k = 0; k < P; ++k) {
    V3<T> pp, pn;
    plane(k, &pp, &pn);
    const bool f0 = dot(p0 - pp, pn) > T(1e-6), f1 = dot(p1 - pp, pn) > T(1e-6);
    const V3<T> cand = seg_plane(p0, p1, pp, pn);
    const V3<T> q0 = f0 ? cand : p0, q1 = f1 ? cand : p1;
    const T v0 = dot(q0 - p0, d01), v1 = dot(q1 - p1, d10);
    if (k == 0 || v0 > b0) { b0 = v0; n0 = q0; }
    if (k == 0 || v1 > b1) { b1 = v1; n1 = q1; }
    both = both || (f0 && f1);
  }
  bool mask = !both;
  if (!mask) { n0 = p0; n1 = p1; }
  if (dot(d10, n0 - n1) < T(0)) mask = false;
  *o0 = n0;
  *o1 = n1;
  return mask;
}

// _clip + _create_contact_manifold :330-449.  cp / sp: clipping (reference) and subject (incident) polygons, P corners each.
template <typename T>
__device__ void contact_manifold(const T* cp, const T* sp, int P, const V3<T>& cnorm, const V3<T>& snorm, T* dist, V3<T>* pos) {
  T inc[4 * MJT_MAXFV * 3];
  unsigned char msk[4 * MJT_MAXFV];
  // subject edges against the clipping polygon's side planes
  for (int j = 0; j < P; ++j) {
    const V3<T> s0 = ldv(sp, j == 0 ? P - 1 : j - 1), s1 = ldv(sp, j);
    V3<T> o0, o1;
    const bool mk = clip_edge_to_planes(s0, s1, P, [&](int k, V3<T>* pp, V3<T>* pn) {
      const V3<T> c0 = ldv(cp, k == 0 ? P - 1 : k - 1), c1 = ldv(cp, k);
      *pp = c0;
      *pn = cross(c1 - c0, cnorm);
    }, &o0, &o1);
    stv(inc, 2 * j, o0);
    stv(inc, 2 * j + 1, o1);
    msk[2 * j] = msk[2 * j + 1] = mk;
  }
  // clipping edges, projected onto the subject plane (:249-257), against the subject polygon's side planes
  const T pd = dot(ldv(sp, 0), snorm);
  const T den = dot(cnorm, snorm);
  const T den_s = den + (den == T(0) ? T(1e-6) : T(0));
  auto project = [&](const V3<T>& p) { return p + cnorm * ((pd - dot(p, snorm)) / den_s); };
  for (int k = 0; k < P; ++k) {
    const V3<T> c0 = project(ldv(cp, k == 0 ? P - 1 : k - 1)), c1 = project(ldv(cp, k));
    V3<T> o0, o1;
    const bool mk = clip_edge_to_planes(c0, c1, P, [&](int j, V3<T>* pp, V3<T>* pn) {
      const V3<T> s0 = ldv(sp, j == 0 ? P - 1 : j - 1), s1 = ldv(sp, j);
      *pp = s0;
      *pn = cross(s1 - s0, snorm);
    }, &o0, &o1);
    stv(inc, 2 * (P + k), o0);
    stv(inc, 2 * (P + k) + 1, o1);
    msk[2 * (P + k)] = msk[2 * (P + k) + 1] = mk;
  }
  const int N = 4 * P;
  const V3<T> nn = normalize3(cnorm), p0 = ldv(cp, 0), ncn = -cnorm;
  for (int i = 0; i < N; ++i) msk[i] = msk[i] && (dot(ldv(inc, i) - p0, ncn) > T(1e-6));
  auto ref_pt = [&](int i) { const V3<T> q = ldv(inc, i); return q - nn * dot(q - p0, nn); };
  int best[4];
  manifold_points<T>(N, ref_pt, [&](int i) { return msk[i] != 0; }, cnorm, best);
  for (int k = 0; k < 4; ++k) {
    const V3<T> c = ref_pt(best[k]);
    const T pen = dot(ldv(inc, best[k]) - c, ncn);
    dist[k] = msk[best[k]] ? -pen : T(1);
    pos[k] = c;
  }
}

template <typename T>
__device__ void plane_convex(const ModelView<T>& m, const EnvBuf<T>& d, int g1, int g2, const int* slot) {
  GeomPose<T> pl, cx;
  load_geom(m, d, g1, pl);
  load_geom(m, d, g2, cx);
  const Hull<T> h = load_hull(m, g2);
  const V3<T> nw = zaxis(pl);
  const V3<T> pp = mat_tmul(cx.mat, pl.pos - cx.pos), n = mat_tmul(cx.mat, nw);
  auto support = [&](int i) { return dot(pp - h.v(i), n); };
  int idx[4];
  manifold_points<T>(h.nvert, [&](int i) { return h.v(i); }, [&](int i) { return support(i) > T(0); }, n, idx);
  T frame[9];
  make_frame(nw, frame);
  for (int k = 0; k < 4; ++k) {
    bool first = true;
    for (int j = 0; j < k; ++j) first = first && (idx[j] != idx[k]);
    const T dist = first ? -support(idx[k]) : T(1);
    write_contact(m, d, slot[k], dist, cx.pos + mat_mul(cx.mat, h.v(idx[k])), frame);
  }
}

// face with the largest negative support (:644-657); support(f) evaluated in face order
template <typename T, typename S>
__device__ __forceinline__ int best_face(int nface, S support, bool* all_negative) {
  int best = 0;
  T bv = T(0);
  bool neg = true;
  for (int f = 0; f < nface; ++f) {
    const T s = support(f);
    neg = neg && (s < T(0));
    const T v = s >= T(0) ? T(-1e12) : s;
    if (f == 0 || v > bv) { bv = v; best = f; }
  }
  if (all_negative) *all_negative = neg;
  return best;
}

// sphere against a hull, everything in the hull's frame (sp = sphere centre): distance, contact position, normal (sphere -> hull)
template <typename T>
__device__ void sphere_hull_core(const Hull<T>& h, const V3<T>& sp, T r, T* dist, V3<T>* pos, V3<T>* nrm) {
  const int f = best_face<T>(h.nface, [&](int k) { const V3<T> nk = h.n(k); return dot(sp - nk * r - h.v(h.face[k * h.nfv]), nk); }, nullptr);
  const V3<T> normal = h.n(f), f0 = h.v(h.face[f * h.nfv]);
  V3<T> pt = sp - normal * dot(sp - f0, normal);
  bool inside = true;
  int ei = 0;
  T ev = T(0);
  for (int k = 0; k < h.nfv; ++k) {
    const V3<T> e0 = h.v(h.face[f * h.nfv + (k == 0 ? h.nfv - 1 : k - 1)]), e1 = h.v(h.face[f * h.nfv + k]);
    const V3<T> en = cross(e1 - e0, normal);
    const T ed = dot(pt - e0, en);
    inside = inside && (ed <= T(0));
    const bool degenerate = en.x == T(0) && en.y == T(0) && en.z == T(0);
    const T v = (degenerate || ed < T(0)) ? T(1e12) : ed;
    if (k == 0 || v < ev) { ev = v; ei = k; }
  }
  if (!inside) {
    const V3<T> e0 = h.v(h.face[f * h.nfv + (ei == 0 ? h.nfv - 1 : ei - 1)]), e1 = h.v(h.face[f * h.nfv + ei]);
    pt = closest_segment_point(e0, e1, pt);
  }
  T dd;
  const V3<T> n = normalize3(pt - sp, &dd);
  const V3<T> spt = sp + n * r;
  *dist = dd - r;
  *pos = (pt + spt) * T(0.5);
  *nrm = n;
}

template <typename T>
__device__ void sphere_convex(const ModelView<T>& m, const EnvBuf<T>& d, int g1, int g2, const int* slot) {
  GeomPose<T> s, cx;
  load_geom(m, d, g1, s);
  load_geom(m, d, g2, cx);
  const Hull<T> h = load_hull(m, g2);
  T dist;
  V3<T> pos, n;
  sphere_hull_core<T>(h, mat_tmul(cx.mat, s.pos - cx.pos), m.F(F_geom_size)[3 * g1], &dist, &pos, &n);
  T frame[9];
  make_frame(mat_mul(cx.mat, n), frame);
  write_contact(m, d, slot[0], dist, mat_mul(cx.mat, pos) + cx.pos, frame);
}

// capsule (segment c0-c1, radius r) against a hull in the hull's frame: two slots with penetration, position and the face / edge
// normal `nrm` (hull -> capsule; the contact normal is its negative)
template <typename T>
__device__ void capsule_hull_core(const Hull<T>& h, const V3<T>& c0, const V3<T>& c1, T r, T* pen, V3<T>* pos, V3<T>* nrm, bool need_support = true) {
  bool has_support;
  const int f = best_face<T>(h.nface, [&](int k) {
    const V3<T> nk = h.n(k), nr = nk * r, fk = h.v(h.face[k * h.nfv]);
    return t_min(dot(c0 - nr - fk, nk), dot(c1 - nr - fk, nk));
  }, &has_support);
  const V3<T> normal = h.n(f), f0 = h.v(h.face[f * h.nfv]);
  auto corner = [&](int k) { return h.v(h.face[f * h.nfv + (k < 0 ? h.nfv - 1 : k)]); };
  V3<T> cl[2];
  const bool mask = clip_edge_to_planes(c0, c1, h.nfv, [&](int k, V3<T>* pp, V3<T>* pn) {
    const V3<T> e0 = corner(k - 1), e1 = corner(k);
    *pp = e0;
    *pn = cross(e1 - e0, normal);
  }, &cl[0], &cl[1]);
  for (int k = 0; k < 2; ++k) {
    const V3<T> c = cl[k] - normal * r;
    const V3<T> fp = c - normal * dot(c - f0, normal);
    pos[k] = (c + fp) * T(0.5);
    nrm[k] = normal;
    // capsule_convex :741-747 also requires every face to see the capsule behind it; _capsule_prism (collision_hfield.py:133-141) does not
    pen[k] = (mask && (has_support || !need_support)) ? dot(fp - c, normal) : T(-1);
  }
  // closest face edge to the capsule axis (:760-784)
  V3<T> ec(0, 0, 0), cc(0, 0, 0);
  T bd = T(0);
  for (int k = 0; k < h.nfv; ++k) {
    V3<T> a, b;
    closest_segment_to_segment(corner(k - 1), corner(k), c0, c1, &a, &b);
    const V3<T> q = a - b;
    const T v = dot(q, q);
    if (k == 0 || v < bd) { bd = v; ec = a; cc = b; }
  }
  T edist;
  const V3<T> eaxis = normalize3(cc - ec, &edist);
  if (r - edist > T(0)) {
    pos[0] = (ec + (cc - eaxis * r)) * T(0.5);
    nrm[0] = eaxis;
    pen[0] = r - edist;
  }
}

template <typename T>
__device__ void capsule_convex(const ModelView<T>& m, const EnvBuf<T>& d, int g1, int g2, const int* slot) {
  GeomPose<T> cap, cx;
  load_geom(m, d, g1, cap);
  load_geom(m, d, g2, cx);
  const Hull<T> h = load_hull(m, g2);
  const T r = m.F(F_geom_size)[3 * g1], len = m.F(F_geom_size)[3 * g1 + 1];
  const V3<T> cpos = mat_tmul(cx.mat, cap.pos - cx.pos), seg = mat_tmul(cx.mat, zaxis(cap)) * len;
  V3<T> pos[2], nrm[2];
  T pen[2];
  capsule_hull_core<T>(h, cpos - seg, cpos + seg, r, pen, pos, nrm);
  for (int k = 0; k < 2; ++k) {
    T frame[9];
    make_frame(mat_mul(cx.mat, -nrm[k]), frame);
    write_contact(m, d, slot[k], -pen[k], cx.pos + mat_mul(cx.mat, pos[k]), frame);
  }
}

// hull A (pose A) against hull B (pose Bp): four manifold slots (distance, position in B's frame) and the separating axis in B's
// frame, oriented A -> B like the reference's _sat_hull_hull (the callers order the hulls: fewer vertices first)
template <typename T>
__device__ void hull_hull_core(const Hull<T>& ha, const GeomPose<T>& A, const Hull<T>& hb, const GeomPose<T>& Bp, T* dist, V3<T>* pos, V3<T>* axis_out) {
  const int P = ha.nfv > hb.nfv ? ha.nfv : hb.nfv;
  // hull A expressed in hull B's frame
  T tm[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) tm[3 * i + j] = Bp.mat[i] * A.mat[j] + Bp.mat[3 + i] * A.mat[3 + j] + Bp.mat[6 + i] * A.mat[6 + j];
  const V3<T> tp = mat_tmul(Bp.mat, A.pos - Bp.pos);
  T va[MJT_VCACHE * 3];
  const int ncache = ha.nvert < MJT_VCACHE ? ha.nvert : MJT_VCACHE;
  for (int i = 0; i < ncache; ++i) stv(va, i, tp + mat_mul(tm, ha.v(i)));
  auto VA = [&](int i) { return i < MJT_VCACHE ? ldv(va, i) : tp + mat_mul(tm, ha.v(i)); };
  auto NA = [&](int f) { return mat_mul(tm, ha.n(f)); };

  // separating-axis scan (:481-530): face normals of A, of B, then edge x edge; first minimum wins.  Axes are projected NB at
  // a time so that every vertex is loaded once per block instead of once per axis (each dot product and each min / max chain is
  // the same arithmetic in the same order as the one-axis form, and the winners are compared in axis order).
  int best_k = 0, best_sign = 1;
  T best_v = T(0);
  V3<T> best_axis(0, 0, 0);
  bool first_axis = true;
  auto take = [&](const V3<T>& ax, int k, T amax, T amin, T bmax, T bmin) {
    const T d1 = amax - bmin, d2 = bmax - amin;
    T sup = t_min(d1, d2);
    if (ax.x == T(0) && ax.y == T(0) && ax.z == T(0)) sup = T(1e6);
    if (first_axis || sup < best_v || (sup == best_v && k < best_k)) { best_v = sup; best_k = k; best_sign = d1 > d2 ? -1 : 1; best_axis = ax; }
    first_axis = false;
  };
  constexpr int NB = 4;
  auto test_axes = [&](const V3<T>* ax, const int* k, int n) {   // n <= NB axes, in scan order
    T amax[NB], amin[NB], bmax[NB], bmin[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) { amax[j] = bmax[j] = -t_inf<T>(); amin[j] = bmin[j] = t_inf<T>(); }
    for (int i = 0; i < ha.nvert; ++i) {
      const V3<T> v = VA(i);
#pragma unroll
      for (int j = 0; j < NB; ++j) { const T s = dot(ax[j], v); amax[j] = t_max(amax[j], s); amin[j] = t_min(amin[j], s); }
    }
    for (int i = 0; i < hb.nvert; ++i) {
      const V3<T> v = hb.v(i);
#pragma unroll
      for (int j = 0; j < NB; ++j) { const T s = dot(ax[j], v); bmax[j] = t_max(bmax[j], s); bmin[j] = t_min(bmin[j], s); }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (j < n) take(ax[j], k[j], amax[j], amin[j], bmax[j], bmin[j]);
  };
  const int nfaces = ha.nface + hb.nface;
  {
    V3<T> ax[NB];
    int kk[NB];
    int n = 0;
    auto push = [&](const V3<T>& a, int k) {
      ax[n] = a;
      kk[n] = k;
      if (++n == NB) { test_axes(ax, kk, NB); n = 0; }
    };
    auto flush = [&]() {
      if (n) {
        for (int j = n; j < NB; ++j) { ax[j] = ax[0]; kk[j] = kk[0]; }   // padding lanes are computed and ignored
        test_axes(ax, kk, n);
        n = 0;
      }
    };
    for (int f = 0; f < ha.nface; ++f) push(NA(f), f);
    for (int f = 0; f < hb.nface; ++f) push(hb.n(f), ha.nface + f);
    for (int ia = 0; ia < ha.nedge; ++ia) {
      const V3<T> da = VA(ha.edge[2 * ia]) - VA(ha.edge[2 * ia + 1]);
      for (int ib = 0; ib < hb.nedge; ++ib) {
        const V3<T> db = hb.v(hb.edge[2 * ib]) - hb.v(hb.edge[2 * ib + 1]);
        push(normalize3(cross(da, db)), nfaces + ib * ha.nedge + ia);
      }
    }
    flush();
  }
  const bool is_edge = best_k >= nfaces;
  // reference / incident faces (:532-560)
  int a_max = 0, a_min = 0, b_max = 0, b_min = 0;
  {
    T vmax = T(0), vmin = T(0);
    for (int f = 0; f < ha.nface; ++f) {
      const T v = dot(best_axis, NA(f));
      if (f == 0 || v > vmax) { vmax = v; a_max = f; }
      if (f == 0 || v < vmin) { vmin = v; a_min = f; }
    }
    for (int f = 0; f < hb.nface; ++f) {
      const T v = dot(best_axis, hb.n(f));
      if (f == 0 || v > vmax) { vmax = v; b_max = f; }
      if (f == 0 || v < vmin) { vmin = v; b_min = f; }
    }
  }
  T ref[MJT_MAXFV * 3], inc[MJT_MAXFV * 3];
  V3<T> ref_n, inc_n;
  if (best_sign > 0) {
    for (int k = 0; k < P; ++k) { stv(ref, k, VA(ha.fv(a_max, k))); stv(inc, k, hb.v(hb.fv(b_min, k))); }
    ref_n = NA(a_max);
    inc_n = hb.n(b_min);
  } else {
    for (int k = 0; k < P; ++k) { stv(ref, k, hb.v(hb.fv(b_max, k))); stv(inc, k, VA(ha.fv(a_min, k))); }
    ref_n = hb.n(b_max);
    inc_n = NA(a_min);
  }
  contact_manifold<T>(ref, inc, P, ref_n, inc_n, dist, pos);
  if (is_edge) {
    int k0 = 0;
    for (int k = 1; k < 4; ++k) if (dist[k] < dist[k0]) k0 = k;
    const T d0 = dist[k0];
    const V3<T> p0 = pos[k0];
    for (int k = 0; k < 4; ++k) { dist[k] = k == 0 ? d0 : T(1); pos[k] = p0; }
  }
  *axis_out = best_axis * T(best_sign);
}

template <typename T>
__device__ void convex_convex(const ModelView<T>& m, const EnvBuf<T>& d, int g1, int g2, const int* slot) {
  Hull<T> ha = load_hull(m, g1), hb = load_hull(m, g2);
  const bool swapped = ha.nvert > hb.nvert;
  if (swapped) { const Hull<T> t = ha; ha = hb; hb = t; const int g = g1; g1 = g2; g2 = g; }
  GeomPose<T> A, Bp;
  load_geom(m, d, g1, A);
  load_geom(m, d, g2, Bp);
  T dist[4];
  V3<T> pos[4], axis;
  hull_hull_core<T>(ha, A, hb, Bp, dist, pos, &axis);
  V3<T> nw = mat_mul(Bp.mat, axis);
  if (swapped) nw = -nw;
  T frame[9];
  make_frame(nw, frame);
  for (int k = 0; k < 4; ++k) write_contact(m, d, slot[k], dist[k], Bp.pos + mat_mul(Bp.mat, pos[k]), frame);
}

template <typename T>
__device__ void convex_pair(const ModelView<T>& m, const EnvBuf<T>& d, int fn, int g1, int g2, const int* slot) {
  switch (fn) {
    case FN_PLANE_CONVEX: plane_convex<T>(m, d, g1, g2, slot); break;
    case FN_SPHERE_CONVEX: sphere_convex<T>(m, d, g1, g2, slot); break;
    case FN_CAPSULE_CONVEX: capsule_convex<T>(m, d, g1, g2, slot); break;
    case FN_CONVEX_CONVEX: convex_convex<T>(m, d, g1, g2, slot); break;
    default: break;
  }
}
