// k_constraint: one thread per (constraint source, env).  A "source" is one equality / friction-loss dof or
// tendon / joint or tendon limit / contact; it owns a fixed, statically known range of efc rows
// (reference constraint.make_constraint :600-768 builds the same rows with per-type vmaps and a concat).
// Each thread writes its dense rows of efc_J plus efc_D, efc_aref, efc_frictionloss.
// Body-point Jacobians use the static ancestor-dof bitmask instead of the reference's reverse tree scan
// (support.jac :138-153).
#pragma once
#include "mjt_common.cuh"
#include "k_smooth.cuh"

#define MJT_MAXROWS 10

// reference constraint._kbi :69-113
template <typename T>
__device__ __forceinline__ void kbi(const ModelView<T>& m, const T* solref, const T* solimp, T pos, T* k, T* b, T* imp) {
  T timeconst = solref[0], dampratio = solref[1];
  if (!(m.h(H_disableflags) & DSBL_REFSAFE)) timeconst = t_max(timeconst, T(2) * m.opt(OPT_timestep)) * (timeconst > T(0) ? T(1) : T(0));
  const T dmin = t_clamp(solimp[0], T(MJT_MINIMP), T(MJT_MAXIMP));
  const T dmax = t_clamp(solimp[1], T(MJT_MINIMP), T(MJT_MAXIMP));
  const T width = t_max(T(MJT_MINVAL), solimp[2]);
  const T mid = t_clamp(solimp[3], T(MJT_MINIMP), T(MJT_MAXIMP));
  const T power = t_max(T(1), solimp[4]);
  T kk = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  T bb = T(2) / (dmax * timeconst);
  if (dampratio <= T(0)) kk = -dampratio / (dmax * dmax);
  if (timeconst <= T(0)) bb = -timeconst / dmax;
  const T imp_x = t_abs(pos) / width;
  // only the branch that is selected is evaluated (each costs two pow() calls); same expressions as the reference's imp_a / imp_b
  T imp_y;
  if (imp_x < mid) imp_y = (T(1) / t_pow(mid, power - T(1))) * t_pow(imp_x, power);
  else imp_y = T(1) - (T(1) / t_pow(T(1) - mid, power - T(1))) * t_pow(T(1) - imp_x, power);
  T im = dmin + imp_y * (dmax - dmin);
  im = t_min(t_max(im, dmin), dmax);
  if (imp_x > T(1)) im = dmax;
  *k = kk; *b = bb; *imp = im;
}

// Jacobian writer of one source.  Full layout: dense efc_J rows (every column written, zeros included).  State layout: efc_J
// is stored compacted to the source's static column support (model.py _row_supports; the form k_solve keeps in shared
// memory): entry (row, support position k) at row_jadr[row] + k.  Columns outside the support are exactly zero by construction
// and are not stored.  Writers visit the columns of a row in ascending order, so the support position is a running cursor.
template <typename T>
struct RowOut {
  const ModelView<T>& m;
  const EnvBuf<T>& d;
  int nv;
  bool compact;
  const int* col;   // support columns of this source (ascending)
  int ns, row0, jadr0;
  int cur;          // cursor into col
  __device__ __forceinline__ void begin(int s) {
    compact = m.h(H_layout_state) != 0;
    col = m.I(I_sup_col) + m.I(I_src_cadr)[s];
    ns = m.I(I_src_nsup)[s];
    row0 = m.I(I_src_row)[s];
    jadr0 = m.I(I_row_jadr)[row0];
    cur = 0;
  }
  // true when column i has to be produced (dense: always; compact: only support columns); positions the cursor on it
  __device__ __forceinline__ bool want(int i) {
    if (!compact) return true;
    while (cur < ns && col[cur] < i) ++cur;
    return cur < ns && col[cur] == i;
  }
  __device__ __forceinline__ void J(int row, int i, T v) const {
    if (compact) d.set(m.D(D_efc_J) + jadr0 + (row - row0) * ns + cur, v);
    else d.set(m.D(D_efc_J) + row * nv + i, v);
  }
  // random-access variant for the few writers that do not walk the columns in order
  __device__ __forceinline__ void Jat(int row, int i, T v) {
    if (compact) {
      for (int k = 0; k < ns; ++k)
        if (col[k] == i) d.set(m.D(D_efc_J) + jadr0 + (row - row0) * ns + k, v);
    } else {
      d.set(m.D(D_efc_J) + row * nv + i, v);
    }
  }
  __device__ __forceinline__ void zero(int r0, int nrows) const {
    if (compact) {
      for (int k = 0; k < nrows * ns; ++k) d.set(m.D(D_efc_J) + jadr0 + (r0 - row0) * ns + k, T(0));
      return;
    }
    for (int r = r0; r < r0 + nrows; ++r)
      for (int i = 0; i < nv; ++i) d.set(m.D(D_efc_J) + r * nv + i, T(0));
  }
  // finish one row: impedance, regulariser, reference acceleration
  __device__ __forceinline__ void finish(int row, T jv, T pos, T pos_norm, T invweight, const T* solref, const T* solimp, T floss) const {
    T k, b, imp;
    kbi(m, solref, solimp, pos_norm, &k, &b, &imp);
    finish_kbi(row, jv, pos, invweight, k, b, imp, floss);
  }
  // the same with (k, b, imp) already evaluated: the rows of one contact share them (same solref / solimp / penetration), and
  // _kbi costs four pow() calls
  __device__ __forceinline__ void finish_kbi(int row, T jv, T pos, T invweight, T k, T b, T imp, T floss) const {
    const T r = t_max(invweight * (T(1) - imp) / imp, T(MJT_MINVAL));
    d.set(m.D(D_efc_D) + row, T(1) / r);
    d.set(m.D(D_efc_aref) + row, -b * jv - k * imp * pos);
    if (m.D(D_efc_frictionloss) >= 0) d.set(m.D(D_efc_frictionloss) + row, floss);
  }
};

template <typename T>
__device__ __forceinline__ bool dof_in_body(const ModelView<T>& m, int body, int i) {
  return (m.I(I_body_dofmask)[body * m.h(H_maskwords) + (i >> 5)] >> (i & 31)) & 1;
}

// jacp_i, jacr_i of world point `pt` on `body` for dof i (zero when i is not an ancestor dof)
template <typename T>
__device__ __forceinline__ void dof_jac(const ModelView<T>& m, const EnvBuf<T>& d, int i, const V3<T>& off, V3<T>* jp, V3<T>* jr) {
  V3<T> a, l;
  ld3(d, m.D(D_cdof) + 6 * i, a);
  ld3(d, m.D(D_cdof) + 6 * i + 3, l);
  *jp = l + cross(a, off);
  *jr = a;
}

template <typename T>
__device__ void constraint_source(const ModelView<T>& m, const MjtIn& in, const EnvBuf<T>& d, size_t B, int env, int s) {
  const int type = m.I(I_src_type)[s], id = m.I(I_src_id)[s], row0 = m.I(I_src_row)[s];
  const int nv = m.h(H_nv);
  RowOut<T> out{m, d, nv};
  out.begin(s);
  const int oQpos = m.D(D_qpos), oQvel = m.D(D_qvel), oCom = m.D(D_subtree_com);
  const int* body_root = m.I(I_body_rootid);
  const T* dof_iw = m.F(F_dof_invweight0);

  bool live = true;   // false when every row of this source is exactly zero this step (J, aref) -- see k_solve's live tile
  switch (type) {
    case SRC_LIM_SH: {  // constraint.py:338-372
      const int qa = m.I(I_jnt_qposadr)[id], da = m.I(I_jnt_dofadr)[id];
      const T q = d.get(oQpos + qa);
      const T dist_min = q - m.F(F_jnt_range)[2 * id], dist_max = m.F(F_jnt_range)[2 * id + 1] - q;
      const T val = dist_min < dist_max ? T(1) : T(-1);
      T pos = t_min(dist_min, dist_max) - m.F(F_jnt_margin)[id];
      const T act = pos < T(0) ? T(1) : T(0);
      live = act != T(0);
      pos *= act;
      out.zero(row0, 1);
      out.Jat(row0, da, val * act);
      out.finish(row0, val * act * d.get(oQvel + da), pos, pos, dof_iw[da], m.F(F_jnt_solref) + 2 * id, m.F(F_jnt_solimp) + 5 * id, T(0));
    } break;
    case SRC_LIM_BALL: {  // :302-335
      const int qa = m.I(I_jnt_qposadr)[id], da = m.I(I_jnt_dofadr)[id];
      Q4<T> q(d.get(oQpos + qa), d.get(oQpos + qa + 1), d.get(oQpos + qa + 2), d.get(oQpos + qa + 3));
      V3<T> axis;
      T angle;
      q2axisangle(q, &axis, &angle);
      T pos = t_max(m.F(F_jnt_range)[2 * id], m.F(F_jnt_range)[2 * id + 1]) - angle - m.F(F_jnt_margin)[id];
      const T act = pos < T(0) ? T(1) : T(0);
      live = act != T(0);
      pos *= act;
      out.zero(row0, 1);
      const T j3[3] = {-axis.x * act, -axis.y * act, -axis.z * act};
      T jv = 0;
      for (int k = 0; k < 3; ++k) { out.Jat(row0, da + k, j3[k]); jv += j3[k] * d.get(oQvel + da + k); }
      out.finish(row0, jv, pos, pos, dof_iw[da], m.F(F_jnt_solref) + 2 * id, m.F(F_jnt_solimp) + 5 * id, T(0));
    } break;
    case SRC_LIM_TEN: {  // :375-405
      const T len = d.get(m.D(D_ten_length) + id);
      const T dist_min = len - m.F(F_ten_range)[2 * id], dist_max = m.F(F_ten_range)[2 * id + 1] - len;
      T pos = t_min(dist_min, dist_max) - m.F(F_ten_margin)[id];
      const T act = pos < T(0) ? T(1) : T(0);
      live = act != T(0);
      const T sign = (dist_min < dist_max ? T(1) : T(-1)) * act;
      pos *= act;
      T jv = 0;
      for (int i = 0; i < nv; ++i) {
        if (!out.want(i)) continue;
        const T v = d.get(m.D(D_ten_J) + id * nv + i) * sign;
        out.J(row0, i, v);
        jv += v * d.get(oQvel + i);
      }
      out.finish(row0, jv, pos, pos, m.F(F_ten_invweight0)[id], m.F(F_ten_solref_lim) + 2 * id, m.F(F_ten_solimp_lim) + 5 * id, T(0));
    } break;
    case SRC_FRIC_DOF: {  // :215-251
      out.zero(row0, 1);
      out.Jat(row0, id, T(1));
      out.finish(row0, d.get(oQvel + id), T(0), T(0), dof_iw[id], m.F(F_dof_solref) + 2 * id, m.F(F_dof_solimp) + 5 * id, m.F(F_dof_frictionloss)[id]);
    } break;
    case SRC_FRIC_TEN: {
      T jv = 0;
      for (int i = 0; i < nv; ++i) {
        if (!out.want(i)) continue;
        const T v = d.get(m.D(D_ten_J) + id * nv + i);
        out.J(row0, i, v);
        jv += v * d.get(oQvel + i);
      }
      out.finish(row0, jv, T(0), T(0), m.F(F_ten_invweight0)[id], m.F(F_ten_solref_fri) + 2 * id, m.F(F_ten_solimp_fri) + 5 * id,
                 m.F(F_ten_frictionloss)[id]);
    } break;
    case SRC_EQ_JOINT: {  // :254-299 (including the reference's wrap-around indexing when joint2 is absent)
      const int j1 = m.I(I_eq_obj1)[id], j2raw = m.I(I_eq_obj2)[id], njnt = m.h(H_njnt);
      const int j2 = j2raw < 0 ? njnt + j2raw : j2raw;
      const T has2 = j2raw > -1 ? T(1) : T(0);
      const T* data = m.F(F_eq_data) + 11 * id;
      const T active = in.eq_active ? T(in.eq_active[(size_t)id * B + env]) : T(1);
      live = active != T(0);
      const int qa1 = m.I(I_jnt_qposadr)[j1], da1 = m.I(I_jnt_dofadr)[j1], qa2 = m.I(I_jnt_qposadr)[j2], da2 = m.I(I_jnt_dofadr)[j2];
      const T pos1 = d.get(oQpos + qa1), pos2 = d.get(oQpos + qa2) * has2;
      const T ref1 = m.F(F_qpos0)[qa1], ref2 = m.F(F_qpos0)[qa2] * has2;
      const T dif = pos2 - ref2;
      T pw[5];
      pw[0] = T(1);
      for (int k = 1; k < 5; ++k) pw[k] = pw[k - 1] * dif;
      T deriv = 0, poly = 0;
      for (int k = 0; k < 4; ++k) deriv += data[1 + k] * pw[k] * T(k + 1);
      for (int k = 0; k < 5; ++k) poly += data[k] * pw[k];
      out.zero(row0, 1);
      T jv = 0;
      out.Jat(row0, da1, T(1) * active);
      out.Jat(row0, da2, -deriv * active);  // same dof: the second write wins, as in the reference's scatter order
      jv = (da1 == da2) ? (-deriv * active * d.get(oQvel + da2)) : (active * d.get(oQvel + da1) - deriv * active * d.get(oQvel + da2));
      const T pos = (pos1 - ref1 - poly) * active;
      out.finish(row0, jv, pos, pos, dof_iw[da1] + dof_iw[da2] * has2, m.F(F_eq_solref) + 2 * id, m.F(F_eq_solimp) + 5 * id, T(0));
    } break;
    case SRC_EQ_CONNECT:
    case SRC_EQ_WELD: {  // :116-212
      const bool weld = type == SRC_EQ_WELD;
      const int b1 = m.I(I_eq_obj1)[id], b2 = m.I(I_eq_obj2)[id];
      const T* data = m.F(F_eq_data) + 11 * id;
      const T active = in.eq_active ? T(in.eq_active[(size_t)id * B + env]) : T(1);
      live = active != T(0);
      const V3<T> anchor1(data[0], data[1], data[2]), anchor2(data[3], data[4], data[5]);
      V3<T> x1, x2;
      T R1[9], R2[9];
      ld3(d, m.D(D_xpos) + 3 * b1, x1);
      ld3(d, m.D(D_xpos) + 3 * b2, x2);
      for (int k = 0; k < 9; ++k) { R1[k] = d.get(m.D(D_xmat) + 9 * b1 + k); R2[k] = d.get(m.D(D_xmat) + 9 * b2 + k); }
      const V3<T> la = weld ? anchor2 : anchor1, lb = weld ? anchor1 : anchor2;
      const V3<T> pos1(R1[0] * la.x + R1[1] * la.y + R1[2] * la.z + x1.x, R1[3] * la.x + R1[4] * la.y + R1[5] * la.z + x1.y,
                       R1[6] * la.x + R1[7] * la.y + R1[8] * la.z + x1.z);
      const V3<T> pos2(R2[0] * lb.x + R2[1] * lb.y + R2[2] * lb.z + x2.x, R2[3] * lb.x + R2[4] * lb.y + R2[5] * lb.z + x2.y,
                       R2[6] * lb.x + R2[7] * lb.y + R2[8] * lb.z + x2.z);
      const V3<T> cpos = pos1 - pos2;
      V3<T> c1, c2;
      ld3(d, oCom + 3 * body_root[b1], c1);
      ld3(d, oCom + 3 * body_root[b2], c2);
      const V3<T> off1 = pos1 - c1, off2 = pos2 - c2;
      const int nrows = weld ? 6 : 3;
      T jv[6] = {0, 0, 0, 0, 0, 0};
      Q4<T> quat, quat1;
      T torquescale = T(1);
      V3<T> crot(0, 0, 0);
      if (weld) {
        torquescale = data[10];
        Q4<T> q1(d.get(m.D(D_xquat) + 4 * b1), d.get(m.D(D_xquat) + 4 * b1 + 1), d.get(m.D(D_xquat) + 4 * b1 + 2), d.get(m.D(D_xquat) + 4 * b1 + 3));
        Q4<T> q2(d.get(m.D(D_xquat) + 4 * b2), d.get(m.D(D_xquat) + 4 * b2 + 1), d.get(m.D(D_xquat) + 4 * b2 + 2), d.get(m.D(D_xquat) + 4 * b2 + 3));
        quat = qmul(q1, Q4<T>(data[6], data[7], data[8], data[9]));
        quat1 = qconj(q2);
        const Q4<T> cr = qmul(quat1, quat);
        crot = V3<T>(cr.x, cr.y, cr.z);
      }
      for (int i = 0; i < nv; ++i) {
        if (!out.want(i)) continue;
        const bool in1 = dof_in_body(m, b1, i), in2 = dof_in_body(m, b2, i);
        V3<T> p(0, 0, 0), r(0, 0, 0);
        if (in1 || in2) {
          V3<T> jp1(0, 0, 0), jr1(0, 0, 0), jp2(0, 0, 0), jr2(0, 0, 0);
          if (in1) dof_jac(m, d, i, off1, &jp1, &jr1);
          if (in2) dof_jac(m, d, i, off2, &jp2, &jr2);
          p = jp1 - jp2;
          r = (jr1 - jr2) * torquescale;
        }
        const T qv = d.get(oQvel + i);
        const T pv[3] = {p.x * active, p.y * active, p.z * active};
        for (int k = 0; k < 3; ++k) { out.J(row0 + k, i, pv[k]); jv[k] += pv[k] * qv; }
        if (weld) {
          // 0.5 * (quat1 * (0, r) * quat).xyz   (constraint.py:187-188)
          const Q4<T> t1(-quat1.x * r.x - quat1.y * r.y - quat1.z * r.z, quat1.w * r.x + quat1.y * r.z - quat1.z * r.y,
                         quat1.w * r.y + quat1.z * r.x - quat1.x * r.z, quat1.w * r.z + quat1.x * r.y - quat1.y * r.x);
          const Q4<T> t2 = qmul(t1, quat);
          const T rv[3] = {T(0.5) * t2.x * active, T(0.5) * t2.y * active, T(0.5) * t2.z * active};
          for (int k = 0; k < 3; ++k) { out.J(row0 + 3 + k, i, rv[k]); jv[3 + k] += rv[k] * qv; }
        }
      }
      T pos[6] = {cpos.x, cpos.y, cpos.z, crot.x * torquescale, crot.y * torquescale, crot.z * torquescale};
      T nrm = 0;
      for (int k = 0; k < nrows; ++k) nrm += pos[k] * pos[k];
      nrm = t_sqrt(nrm) * active;
      const T* iw1 = m.F(F_body_invweight0) + 2 * b1;
      const T* iw2 = m.F(F_body_invweight0) + 2 * b2;
      T kk, bb, imp;
      kbi(m, m.F(F_eq_solref) + 2 * id, m.F(F_eq_solimp) + 5 * id, nrm, &kk, &bb, &imp);   // one evaluation for the 3 / 6 rows
      for (int k = 0; k < nrows; ++k) {
        const T iw = (k < 3) ? (iw1[0] + iw2[0]) : (iw1[1] + iw2[1]);
        out.finish_kbi(row0 + k, jv[k], pos[k] * active, iw, kk, bb, imp, T(0));
      }
    } break;
    case SRC_CONTACT: {  // :408-583
      const int c = id;                                    // contact slot: where dist / pos / frame live
      // parameter index: the slot itself, or with max_contact_points the candidate k_topk put into this slot
      const int cc = m.h(H_topk) > 0 ? (int)d.get(m.D(D_scr_consrc) + c) : c;
      const int condim = m.I(I_con_dim)[cc], b1 = m.I(I_con_body1)[cc], b2 = m.I(I_con_body2)[cc];
      const bool elliptic = m.h(H_cone) == 1;
      const int nrows = condim == 1 ? 1 : (elliptic ? condim : 2 * (condim - 1));
      const T dist = d.get(m.D(D_contact_dist) + c) - m.F(F_con_includemargin)[cc];
      const bool active = dist < T(0);
      live = active;
      const T* fri = m.F(F_con_friction) + 5 * cc;
      const T t = m.F(F_body_invweight0)[2 * b1] + m.F(F_body_invweight0)[2 * b2];
      T jv[MJT_MAXROWS];
      for (int k = 0; k < MJT_MAXROWS; ++k) jv[k] = 0;
      if (!active) {
        out.zero(row0, nrows);
      } else {
        V3<T> pt, c1, c2;
        T fr[9];
        ld3(d, m.D(D_contact_pos) + 3 * c, pt);
        for (int k = 0; k < 9; ++k) fr[k] = d.get(m.D(D_contact_frame) + 9 * c + k);
        ld3(d, oCom + 3 * body_root[b1], c1);
        ld3(d, oCom + 3 * body_root[b2], c2);
        const V3<T> off1 = pt - c1, off2 = pt - c2;
        const V3<T> f0(fr[0], fr[1], fr[2]), f1(fr[3], fr[4], fr[5]), f2(fr[6], fr[7], fr[8]);
        for (int i = 0; i < nv; ++i) {
          if (!out.want(i)) continue;
          const bool in1 = dof_in_body(m, b1, i), in2 = dof_in_body(m, b2, i);
          T dd[6] = {0, 0, 0, 0, 0, 0};
          if (in1 != in2) {  // common ancestors cancel exactly
            V3<T> jp, jr;
            dof_jac(m, d, i, in2 ? off2 : off1, &jp, &jr);
            const T sgn = in2 ? T(1) : T(-1);
            dd[0] = dot(f0, jp) * sgn; dd[1] = dot(f1, jp) * sgn; dd[2] = dot(f2, jp) * sgn;
            if (condim > 3) { dd[3] = dot(f0, jr) * sgn; dd[4] = dot(f1, jr) * sgn; dd[5] = dot(f2, jr) * sgn; }
          }
          const T qv = d.get(oQvel + i);
          if (condim == 1) {
            out.J(row0, i, dd[0]);
            jv[0] += dd[0] * qv;
          } else if (!elliptic) {
            for (int e = 0; e < nrows; ++e) {
              const int k = e >> 1;
              const T v = dd[0] + dd[1 + k] * ((e & 1) ? -fri[k] : fri[k]);
              out.J(row0 + e, i, v);
              jv[e] += v * qv;
            }
          } else {
            for (int e = 0; e < nrows; ++e) {
              out.J(row0 + e, i, dd[e]);
              jv[e] += dd[e] * qv;
            }
          }
        }
      }
      const T* solref = m.F(F_con_solref) + 2 * cc;
      const T* solimp = m.F(F_con_solimp) + 5 * cc;
      const T posa = active ? dist : T(0);
      if (condim == 1) {
        out.finish(row0, jv[0], posa, posa, t, solref, solimp, T(0));
      } else if (!elliptic) {
        const T mu = fri[0];
        const T iw = (t + mu * mu * t) * T(2) * mu * mu / m.opt(OPT_impratio);
        T k, b, imp;
        kbi(m, solref, solimp, posa, &k, &b, &imp);   // identical for every row of the pyramid
        for (int e = 0; e < nrows; ++e) out.finish_kbi(row0 + e, jv[e], posa, iw, k, b, imp, T(0));
      } else {
        const T* srf0 = m.F(F_con_solreffriction) + 2 * cc;
        const bool any = srf0[0] != T(0) || srf0[1] != T(0);
        const T srf[2] = {any ? srf0[0] : srf0[0] + solref[0], any ? srf0[1] : srf0[1] + solref[1]};
        T k0, b0, imp0, k1, b1, imp1;   // normal row / friction rows (the impedance is the same: it depends on dist and solimp only)
        kbi(m, solref, solimp, dist, &k0, &b0, &imp0);
        kbi(m, srf, solimp, dist, &k1, &b1, &imp1);
        for (int e = 0; e < nrows; ++e) {
          T iw = t;
          if (e >= 1) iw = t / m.opt(OPT_impratio);
          if (e >= 2) iw = iw * (fri[0] * fri[0]) / (fri[e - 1] * fri[e - 1]);
          out.finish_kbi(row0 + e, jv[e], e == 0 ? posa : T(0), iw, e == 0 ? k0 : k1, e == 0 ? b0 : b1, e == 0 ? imp0 : imp1, T(0));
        }
      }
    } break;
    default: break;
  }
  d.set(m.D(D_scr_live) + s, live ? T(1) : T(0));
}

template <typename T>
// resident CTAs per SM (register cap): f32 gains from 8 (64 registers; measured 0.232 -> 0.187 ms Humanoid B = 65536), f64 spills there
__global__ void __launch_bounds__(MJT_ENV_THREADS, (sizeof(T) == 4 ? 2 : 1) * MJT_CONSTRAINT_MINB) k_constraint(const int* __restrict__ gblob, int blob_words, MjtIn in, T* __restrict__ out, int B) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  // grid.x walks the row sources (fastest), grid.y the env blocks: all sources of one env block are co-resident, so
  // the kinematic quantities they share (cdof, xpos, contact frames) are read from DRAM once and then hit in L2
  for (int env = blockIdx.y * blockDim.x + threadIdx.x; env < B; env += gridDim.y * blockDim.x) {
    EnvBuf<T> d{out + env, (size_t)B};
    for (int s = blockIdx.x; s < m.h(H_nsrc); s += gridDim.x) constraint_source<T>(m, in, d, (size_t)B, env, s);
  }
}
