// k_hfield: height-field narrowphase, one thread per (pair, env), called from k_collision.
// Follows reference collision_hfield.py: _build_prisms :233-294 (two triangular prisms per grid cell of the sub-grid around the
// object), _sphere_prism :44-91, _capsule_prism :94-168, _convex_prism :171-230, _hfield_collision :297-335 (normals flipped, those
// not pointing up replaced by the prism's top normal), _select_manifold :338-349 (4 contacts out of all prism contacts), and
// mesh.hfield_prism (mesh.py:274-344).  The reference decides in Python, per env, which cells the object covers; here the sub-grid
// SIZE is static (model.py: pair_hf, from geom_rbound and the grid pitch like collision_driver._hfield_subgrid_size :318-329) and
// its origin (cmin, rmin) is computed per env from the object's position -- "the reference, env by env".
// A prism is a 6-vertex hull whose tables live in local memory, so the sphere / capsule / hull cores of k_convex.cuh apply as is.
#pragma once
#include "k_convex.cuh"

#define MJT_HF_MAXCAND 128   // prisms x contacts per prism (model.py HFIELD_MAX_PRISMS = 32, up to 4 contacts each)

// mesh.hfield_prism: vertices 0-2 bottom triangle, 3-5 top triangle; faces padded to 4 corners; the 9 unique edges
__device__ const int kPrismFace[20] = {0, 1, 2, 0, 3, 4, 5, 3, 0, 3, 5, 1, 0, 2, 4, 3, 2, 1, 5, 4};
__device__ const int kPrismEdge[18] = {0, 1, 1, 2, 0, 2, 3, 4, 3, 5, 4, 5, 0, 3, 1, 5, 2, 4};

template <typename T>
struct PrismHull {
  T vert[18], norm[15];
  V3<T> centroid;
  __device__ __forceinline__ Hull<T> hull() const {
    Hull<T> h;
    h.vert = vert; h.norm = norm; h.face = kPrismFace; h.edge = kPrismEdge;
    h.nvert = 6; h.nface = 5; h.nfv = 4; h.nedge = 9;
    return h;
  }
  // top: the three surface points (p1, p2, p3); bottom = the same at z = -zbase, reordered [0, 2, 1] (collision_hfield.py:268-273)
  __device__ void build(const V3<T>* top, T zbase) {
    V3<T> v[6];
    v[0] = V3<T>(top[0].x, top[0].y, -zbase); v[1] = V3<T>(top[2].x, top[2].y, -zbase); v[2] = V3<T>(top[1].x, top[1].y, -zbase);
    v[3] = top[0]; v[4] = top[1]; v[5] = top[2];
    V3<T> c(0, 0, 0);
    for (int i = 0; i < 6; ++i) c = c + v[i];
    c = V3<T>(c.x / T(6), c.y / T(6), c.z / T(6));
    centroid = c;
    for (int i = 0; i < 6; ++i) stv(vert, i, v[i] - c);
    for (int f = 0; f < 5; ++f) {   // ccw winding, the last (repeated) corner unused (mesh.py:319-323)
      const V3<T> a = ldv(vert, kPrismFace[4 * f]), b = ldv(vert, kPrismFace[4 * f + 1]), cc = ldv(vert, kPrismFace[4 * f + 2]);
      stv(norm, f, normalize3(cross(cc - b, a - b)));
    }
  }
};

template <typename T>
__device__ void hfield_pair(const ModelView<T>& m, const EnvBuf<T>& d, int p, int fn, int g1, int g2, const int* slot) {
  const int* ph = m.I(I_pair_hf) + 3 * p;
  const int hid = ph[0], xb = ph[1], yb = ph[2];
  const T rb = m.F(F_pair_rbound)[p];
  const T* hs = m.F(F_hf_size) + 4 * hid;
  const int nrow = m.I(I_hf_nrow)[hid], ncol = m.I(I_hf_ncol)[hid];
  const T* data = m.F(F_hf_data) + m.I(I_hf_adr)[hid];   // data[row * ncol + col] in [0, 1]
  const T xhalf = hs[0], yhalf = hs[1], zscale = hs[2], zbase = hs[3];
  GeomPose<T> H, O;
  load_geom(m, d, g1, H);
  load_geom(m, d, g2, O);
  // the object in the height field's frame (_hfield_collision :309-311)
  GeomPose<T> L;
  L.pos = mat_tmul(H.mat, O.pos - H.pos);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) L.mat[3 * i + j] = H.mat[i] * O.mat[j] + H.mat[3 + i] * O.mat[3 + j] + H.mat[6 + i] * O.mat[6 + j];
  const T dx = T(2) * xhalf / T(ncol - 1), dy = T(2) * yhalf / T(nrow - 1);
  const int cmin = (int)floor((L.pos.x - rb + xhalf) / (T(2) * xhalf) * T(ncol - 1));
  const int rmin = (int)floor((L.pos.y - rb + yhalf) / (T(2) * yhalf) * T(nrow - 1));
  const T* size = m.F(F_geom_size) + 3 * g2;
  const int npp = fn == FN_HFIELD_SPHERE ? 1 : (fn == FN_HFIELD_CAPSULE ? 2 : 4);

  T cd[MJT_HF_MAXCAND], cp[MJT_HF_MAXCAND * 3], cn[MJT_HF_MAXCAND * 3];
  int nc = 0;
  PrismHull<T> pr;
  auto surf = [&](int col, int row) { return V3<T>(dx * T(col) - xhalf, dy * T(row) - yhalf, data[row * ncol + col] * zscale); };
  for (int r = 0; r < yb; ++r)
    for (int c = 0; c < xb; ++c) {
      int ri = rmin + r, ci = cmin + c;
      ri = ri < 0 ? 0 : (ri > nrow - 2 ? nrow - 2 : ri);
      ci = ci < 0 ? 0 : (ci > ncol - 2 ? ncol - 2 : ci);
      const V3<T> p1 = surf(ci, ri), p2 = surf(ci + 1, ri + 1), p3 = surf(ci, ri + 1), p2b = surf(ci + 1, ri);
      for (int half = 0; half < 2 && nc + npp <= MJT_HF_MAXCAND; ++half) {
        V3<T> top[3];
        top[0] = p1;
        top[1] = half == 0 ? p2 : p2b;
        top[2] = half == 0 ? p3 : p2;
        pr.build(top, zbase);
        const Hull<T> h = pr.hull();
        const V3<T> topn = ldv(pr.norm, 1);
        if (fn == FN_HFIELD_SPHERE) {
          T dist;
          V3<T> pos, n;
          sphere_hull_core<T>(h, L.pos - pr.centroid, size[0], &dist, &pos, &n);
          cd[nc] = dist;
          stv(cp, nc, pos + pr.centroid);
          stv(cn, nc, n);
        } else if (fn == FN_HFIELD_CAPSULE) {
          const V3<T> cpos = L.pos - pr.centroid, seg = V3<T>(L.mat[2], L.mat[5], L.mat[8]) * size[1];
          T pen[2];
          V3<T> pos[2], nrm[2];
          capsule_hull_core<T>(h, cpos - seg, cpos + seg, size[0], pen, pos, nrm, false);
          for (int k = 0; k < 2; ++k) { cd[nc + k] = -pen[k]; stv(cp, nc + k, pos[k] + pr.centroid); stv(cn, nc + k, -nrm[k]); }
        } else {
          const Hull<T> ho = load_hull(m, g2);
          GeomPose<T> P;
          P.pos = pr.centroid;
#pragma unroll
          for (int k = 0; k < 9; ++k) P.mat[k] = (k % 4 == 0) ? T(1) : T(0);
          const bool swapped = ho.nvert > 6;   // the hull with fewer vertices comes first (_convex_prism :184-199)
          T dist[4];
          V3<T> pos[4], axis;
          if (swapped) hull_hull_core<T>(h, P, ho, L, dist, pos, &axis);
          else hull_hull_core<T>(ho, L, h, P, dist, pos, &axis);
          const GeomPose<T>& Bp = swapped ? L : P;
          V3<T> nw = mat_mul(Bp.mat, axis);
          if (swapped) nw = -nw;
          for (int k = 0; k < 4; ++k) { cd[nc + k] = dist[k]; stv(cp, nc + k, Bp.pos + mat_mul(Bp.mat, pos[k])); stv(cn, nc + k, nw); }
        }
        // _hfield_collision :321-333: flip, and take the prism's top normal where the flipped normal does not point up
        for (int k = 0; k < npp; ++k) {
          V3<T> n = -ldv(cn, nc + k);
          if (n.z < T(1e-6)) n = topn;
          stv(cn, nc + k, n);
        }
        nc += npp;
      }
    }
  // _select_manifold :338-349
  V3<T> nmean(0, 0, 0);
  T dmin = t_inf<T>();
  for (int i = 0; i < nc; ++i) { nmean = nmean + ldv(cn, i); dmin = t_min(dmin, cd[i]); }
  nmean = V3<T>(nmean.x / T(nc), nmean.y / T(nc), nmean.z / T(nc));
  const T thr = t_min(T(0), dmin + T(1e-3));
  int idx[4];
  manifold_points<T>(nc, [&](int i) { return ldv(cp, i); }, [&](int i) { return cd[i] < thr; }, nmean, idx);
  for (int k = 0; k < 4; ++k) {
    bool first = true;
    for (int j = 0; j < k; ++j) first = first && (idx[j] != idx[k]);
    T frame[9];
    make_frame(mat_mul(H.mat, ldv(cn, idx[k])), frame);
    write_contact(m, d, slot[k], first ? cd[idx[k]] : T(1), mat_mul(H.mat, ldv(cp, idx[k])) + H.pos, frame);
  }
}
