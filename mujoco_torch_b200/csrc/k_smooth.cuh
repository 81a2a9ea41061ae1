// k_smooth: one thread per environment, env-minor coalesced access.
// Fuses the reference's _check_state -> kinematics -> com_pos -> tendon -> crb/make_m -> transmission ->
// com_vel -> passive -> rne -> actuation -> qfrc_smooth (reference forward.py:44-59,73-99,102-228;
// smooth.py:34-308,385-591; passive.py:80-200; support.py:50-80,138-194).
// The reference's tree scans become loops in body order (parents precede children).
#pragma once
#include "mjt_common.cuh"

template <typename T>
__device__ __forceinline__ void ld3(const EnvBuf<T>& d, int off, V3<T>& v) { v.x = d.get(off); v.y = d.get(off + 1); v.z = d.get(off + 2); }
template <typename T>
__device__ __forceinline__ void st3(const EnvBuf<T>& d, int off, const V3<T>& v) { d.set(off, v.x); d.set(off + 1, v.y); d.set(off + 2, v.z); }
template <typename T>
__device__ __forceinline__ V3<T> f3(const T* p) { return V3<T>(p[0], p[1], p[2]); }
template <typename T>
__device__ __forceinline__ Q4<T> f4(const T* p) { return Q4<T>(p[0], p[1], p[2], p[3]); }

// Jacobian-transpose application of a force/torque at a world point on body b:
// qfrc[i] += jacp_i . force + jacr_i . torque over ancestor dofs (reference support.apply_ft :169-181).
template <typename T, typename F>
__device__ __forceinline__ void for_ancestor_dofs(const ModelView<T>& m, int b, F fn) {
  const int* bdofadr = m.I(I_body_dofadr);
  const int* bdofnum = m.I(I_body_dofnum);
  const int* dparent = m.I(I_dof_parentid);
  const int* bparent = m.I(I_body_parentid);
  while (b > 0 && bdofnum[b] == 0) b = bparent[b];
  if (b <= 0) return;
  for (int i = bdofadr[b] + bdofnum[b] - 1; i >= 0; i = dparent[i]) fn(i);
}

template <typename T>
__device__ void smooth_env(const ModelView<T>& m, const MjtIn& in, const EnvBuf<T>& d, size_t B, int env, int mode) {
  const int nq = m.h(H_nq), nv = m.h(H_nv), nu = m.h(H_nu), na = m.h(H_na), nbody = m.h(H_nbody), njnt = m.h(H_njnt);
  const int ngeom = m.h(H_ngeom), nsite = m.h(H_nsite), ntendon = m.h(H_ntendon), flags = m.h(H_disableflags);
  const int oQpos = m.D(D_qpos), oQvel = m.D(D_qvel), oXpos = m.D(D_xpos), oXquat = m.D(D_xquat), oXmat = m.D(D_xmat);
  const int oXipos = m.D(D_xipos), oXimat = m.D(D_ximat), oAnchor = m.D(D_xanchor), oAxis = m.D(D_xaxis);
  const int oCom = m.D(D_subtree_com), oCinert = m.D(D_cinert), oCrb = m.D(D_crb), oCdof = m.D(D_cdof), oCdofdot = m.D(D_cdof_dot);
  const int oCvel = m.D(D_cvel), oM = m.D(D_qM);
  // state layout (layout.py): fields nothing downstream reads have offset -1 and are not produced; qM is a packed lower triangle
  const bool mpacked = m.h(H_layout_state) != 0;
  const bool check = (mode & MJT_MODE_CHECK) != 0;
  const int* body_parent = m.I(I_body_parentid);
  const int* body_root = m.I(I_body_rootid);
  const int* body_jntadr = m.I(I_body_jntadr);
  const int* body_jntnum = m.I(I_body_jntnum);
  const int* body_dofadr = m.I(I_body_dofadr);
  const int* body_dofnum = m.I(I_body_dofnum);
  const int* jnt_type = m.I(I_jnt_type);
  const int* jnt_qposadr = m.I(I_jnt_qposadr);
  const int* jnt_dofadr = m.I(I_jnt_dofadr);
  const int* jnt_bodyid = m.I(I_jnt_bodyid);
  const int* dof_bodyid = m.I(I_dof_bodyid);
  const int* dof_parent = m.I(I_dof_parentid);
  const T* qpos0 = m.F(F_qpos0);

  // ---- _check_state (forward.py:44-59) + copy of pass-through state
  for (int i = 0; i < nq; ++i) {
    T q = in_get<T>(in.qpos, B, env, i);
    if (check && (!t_finite(q) || t_abs(q) > T(MJT_MAXVAL))) q = qpos0[i];
    d.set(oQpos + i, q);
  }
  for (int i = 0; i < nv; ++i) {
    T v = in_get<T>(in.qvel, B, env, i);
    if (check && (!t_finite(v) || t_abs(v) > T(MJT_MAXVAL))) v = T(0);
    d.set(oQvel + i, v);
  }
  for (int i = 0; i < na; ++i) d.set(m.D(D_act) + i, in_get<T>(in.act, B, env, i));
  d.set(m.D(D_time), in_get<T>(in.time, B, env, 0));

  const int* body_mocap = m.I(I_body_mocapid);
  const T* body_mass = m.F(F_body_mass);
  const T* body_submass = m.F(F_body_subtreemass);

  // ---- pass 1, bodies root -> leaves: kinematics (smooth.py:34-207).  Everything that depends only on the body's
  // pose is produced while the pose is in registers: xmat, xipos, ximat, joint anchors/axes, the body's geoms and
  // sites (support.local_to_global :99-108) and the subtree-COM numerator (smooth.py:215-221).
  {
    V3<T> lpos(0, 0, 0);   // tree pose of body b-1 (what its children see)
    Q4<T> lquat(1, 0, 0, 0);
    for (int b = 0; b < nbody; ++b) {
      V3<T> pos = f3(m.F(F_body_pos) + 3 * b);
      Q4<T> quat = f4(m.F(F_body_quat) + 4 * b);
      if (b > 0) {
        const int p = body_parent[b];
        V3<T> ppos;
        Q4<T> pq;
        if (p == b - 1) {
          ppos = lpos;
          pq = lquat;
        } else if (body_mocap[p] >= 0) {   // mocap overrides are not seen by the tree pass (smooth.py:105-113)
          ppos = f3(m.F(F_body_pos) + 3 * p);
          pq = f4(m.F(F_body_quat) + 4 * p);
        } else {
          ld3(d, oXpos + 3 * p, ppos);
          pq = Q4<T>(d.get(oXquat + 4 * p), d.get(oXquat + 4 * p + 1), d.get(oXquat + 4 * p + 2), d.get(oXquat + 4 * p + 3));
        }
        pos = ppos + qrot(pos, pq);
        quat = qmul(pq, quat);
      }
      for (int j = body_jntadr[b], je = j + body_jntnum[b]; j < je; ++j) {
        const int jt = jnt_type[j], qa = jnt_qposadr[j];
        const V3<T> jpos = f3(m.F(F_jnt_pos) + 3 * j), jaxis = f3(m.F(F_jnt_axis) + 3 * j);
        V3<T> anchor, axis;
        if (jt == JNT_FREE) {
          ld3(d, oQpos + qa, anchor);
          axis = V3<T>(0, 0, 1);
        } else {
          anchor = qrot(jpos, quat) + pos;
          axis = qrot(jaxis, quat);
        }
        st3(d, oAnchor + 3 * j, anchor);
        st3(d, oAxis + 3 * j, axis);
        if (jt == JNT_FREE) {
          pos = anchor;
          quat = qnormalize(Q4<T>(d.get(oQpos + qa + 3), d.get(oQpos + qa + 4), d.get(oQpos + qa + 5), d.get(oQpos + qa + 6)));
          d.set(oQpos + qa + 3, quat.w); d.set(oQpos + qa + 4, quat.x); d.set(oQpos + qa + 5, quat.y); d.set(oQpos + qa + 6, quat.z);
        } else if (jt == JNT_BALL) {
          Q4<T> ql = qnormalize(Q4<T>(d.get(oQpos + qa), d.get(oQpos + qa + 1), d.get(oQpos + qa + 2), d.get(oQpos + qa + 3)));
          d.set(oQpos + qa, ql.w); d.set(oQpos + qa + 1, ql.x); d.set(oQpos + qa + 2, ql.y); d.set(oQpos + qa + 3, ql.z);
          quat = qmul(quat, ql);
          pos = anchor - qrot(jpos, quat);
        } else if (jt == JNT_HINGE) {
          const T angle = d.get(oQpos + qa) - qpos0[qa];
          quat = qmul(quat, axis_angle_q(jaxis, angle));
          pos = anchor - qrot(jpos, quat);
        } else {
          pos = pos + axis * (d.get(oQpos + qa) - qpos0[qa]);
        }
      }
      lpos = pos;
      lquat = quat;
      const int mid = body_mocap[b];
      if (mid >= 0) {
        pos = V3<T>(in_get<T>(in.mocap_pos, B, env, 3 * mid), in_get<T>(in.mocap_pos, B, env, 3 * mid + 1), in_get<T>(in.mocap_pos, B, env, 3 * mid + 2));
        quat = qnormalize(Q4<T>(in_get<T>(in.mocap_quat, B, env, 4 * mid), in_get<T>(in.mocap_quat, B, env, 4 * mid + 1),
                                in_get<T>(in.mocap_quat, B, env, 4 * mid + 2), in_get<T>(in.mocap_quat, B, env, 4 * mid + 3)));
      }
      st3(d, oXpos + 3 * b, pos);
      d.set(oXquat + 4 * b, quat.w); d.set(oXquat + 4 * b + 1, quat.x); d.set(oXquat + 4 * b + 2, quat.y); d.set(oXquat + 4 * b + 3, quat.z);
      T mat[9];
      if (oXmat >= 0) {
        q2mat(quat, mat);
#pragma unroll
        for (int k = 0; k < 9; ++k) d.set(oXmat + 9 * b + k, mat[k]);
      }
      const V3<T> ipos = pos + qrot(f3(m.F(F_body_ipos) + 3 * b), quat);
      st3(d, oXipos + 3 * b, ipos);
      if (oXimat >= 0) {
        q2mat(qmul(quat, f4(m.F(F_body_iquat) + 4 * b)), mat);
#pragma unroll
        for (int k = 0; k < 9; ++k) d.set(oXimat + 9 * b + k, mat[k]);
      }
      st3(d, oCom + 3 * b, ipos * body_mass[b]);
      for (int pass = 0; pass < 2; ++pass) {
        const int* list = m.I(pass == 0 ? I_body_geom : I_body_site) + m.I(pass == 0 ? I_body_geomadr : I_body_siteadr)[b];
        const int n = m.I(pass == 0 ? I_body_geomnum : I_body_sitenum)[b];
        const T* lp = m.F(pass == 0 ? F_geom_pos : F_site_pos);
        const T* lq = m.F(pass == 0 ? F_geom_quat : F_site_quat);
        const int oP = m.D(pass == 0 ? D_geom_xpos : D_site_xpos), oR = m.D(pass == 0 ? D_geom_xmat : D_site_xmat);
        if (oP < 0) continue;
        const int* gneed = m.I(I_geom_need);   // state layout: only geoms a collision pair or a sensor reads are placed
        for (int k = 0; k < n; ++k) {
          const int g = list[k];
          if (pass == 0 && !gneed[g]) continue;
          st3(d, oP + 3 * g, pos + qrot(f3(lp + 3 * g), quat));
          q2mat(qmul(quat, f4(lq + 4 * g)), mat);
#pragma unroll
          for (int c = 0; c < 9; ++c) d.set(oR + 9 * g + c, mat[c]);
        }
      }
    }
  }
  // ---- pass 2, leaves -> root: subtree-COM numerators (smooth.py:222-226)
  for (int b = nbody - 1; b > 0; --b) {
    const int p = body_parent[b];
    V3<T> a, c;
    ld3(d, oCom + 3 * p, a);
    ld3(d, oCom + 3 * b, c);
    st3(d, oCom + 3 * p, a + c);
  }

  // ---- pass 3, root -> leaves: subtree_com, cinert, cdof (smooth.py:227-288), com_vel (smooth.py:385-424) and the
  // forward half of rne (smooth.py:427-452).  Each body's spatial quantities are consumed from registers.
  const int oCacc = m.D(D_scr_cacc), oCfrc = m.D(D_scr_cfrc), oBias = m.D(D_qfrc_bias);
  {
    V3<T> rc(0, 0, 0);
    int rc_root = -1;
    T lv[6], la[6];   // cvel / cacc of body b-1
    for (int b = 0; b < nbody; ++b) {
      V3<T> pos;
      ld3(d, oXpos + 3 * b, pos);
      const Q4<T> quat(d.get(oXquat + 4 * b), d.get(oXquat + 4 * b + 1), d.get(oXquat + 4 * b + 2), d.get(oXquat + 4 * b + 3));
      const V3<T> ipos = pos + qrot(f3(m.F(F_body_ipos) + 3 * b), quat);
      V3<T> s;
      ld3(d, oCom + 3 * b, s);
      const T msub = body_submass[b];
      if (msub < T(MJT_MINVAL)) {
        ld3(d, oXipos + 3 * b, s);
      } else {
        const T dm = t_max(msub, T(MJT_MINVAL));
        s = V3<T>(s.x / dm, s.y / dm, s.z / dm);
      }
      st3(d, oCom + 3 * b, s);
      const int root = body_root[b];
      if (root == b) { rc = s; rc_root = b; }
      else if (root != rc_root) { ld3(d, oCom + 3 * root, rc); rc_root = root; }
      // cinert = R diag(I) R^T + m (|o|^2 E - o o^T), [m o, m]
      T ci[10];
      {
        const V3<T> off = ipos - rc;
        T R[9];
        q2mat(qmul(quat, f4(m.F(F_body_iquat) + 4 * b)), R);
        const T* I3 = m.F(F_body_inertia) + 3 * b;
        const T ms = body_mass[b];
        const int ti[6] = {0, 1, 2, 0, 0, 1}, tj[6] = {0, 1, 2, 1, 2, 2};
        const T o[3] = {off.x, off.y, off.z};
        const T oo = dot(off, off);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int a = ti[k], c = tj[k];
          T v = R[3 * a] * I3[0] * R[3 * c] + R[3 * a + 1] * I3[1] * R[3 * c + 1] + R[3 * a + 2] * I3[2] * R[3 * c + 2];
          v += ((a == c ? oo : T(0)) - o[a] * o[c]) * ms;
          ci[k] = v;
        }
        ci[6] = off.x * ms; ci[7] = off.y * ms; ci[8] = off.z * ms; ci[9] = ms;
#pragma unroll
        for (int k = 0; k < 10; ++k) { if (oCinert >= 0) d.set(oCinert + 10 * b + k, ci[k]); d.set(oCrb + 10 * b + k, ci[k]); }
      }
      // parent's cvel / cacc
      T v[6], a[6];
      if (b == 0) {
        const bool nograv = flags & DSBL_GRAVITY;
#pragma unroll
        for (int k = 0; k < 6; ++k) { v[k] = T(0); a[k] = T(0); }
        if (!nograv) { a[3] = -m.opt(OPT_gravity0); a[4] = -m.opt(OPT_gravity1); a[5] = -m.opt(OPT_gravity2); }
      } else {
        const int p = body_parent[b];
        if (p == b - 1) {
#pragma unroll
          for (int k = 0; k < 6; ++k) { v[k] = lv[k]; a[k] = la[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) { v[k] = d.get(oCvel + 6 * p + k); a[k] = d.get(oCacc + 6 * p + k); }
        }
      }
      T xm[9];
      q2mat(quat, xm);
      for (int j = body_jntadr[b], je = j + body_jntnum[b]; j < je; ++j) {
        const int jt = jnt_type[j], da = jnt_dofadr[j];
        V3<T> an, ax;
        ld3(d, oAnchor + 3 * j, an);
        const V3<T> off = rc - an;
        int first = da, cnt = 1;
        T cd[3][6];
        if (jt == JNT_FREE) {
          for (int k = 0; k < 3; ++k) {   // translational dofs: unit twists; their cdof_dot is zero (smooth.py:400-404)
            const T qv = d.get(oQvel + da + k);
            for (int c = 0; c < 6; ++c) {
              const T e = (c == 3 + k) ? T(1) : T(0);
              d.set(oCdof + 6 * (da + k) + c, e);
              if (oCdofdot >= 0) d.set(oCdofdot + 6 * (da + k) + c, T(0));
              v[c] += e * qv;
            }
          }
          first = da + 3;
        }
        if (jt == JNT_FREE || jt == JNT_BALL) {
          cnt = 3;
          for (int k = 0; k < 3; ++k) {   // body axes = columns of xmat
            const V3<T> w(xm[k], xm[3 + k], xm[6 + k]);
            const V3<T> l = cross(w, off);
            cd[k][0] = w.x; cd[k][1] = w.y; cd[k][2] = w.z; cd[k][3] = l.x; cd[k][4] = l.y; cd[k][5] = l.z;
          }
        } else {
          ld3(d, oAxis + 3 * j, ax);
          if (jt == JNT_HINGE) {
            const V3<T> l = cross(ax, off);
            cd[0][0] = ax.x; cd[0][1] = ax.y; cd[0][2] = ax.z; cd[0][3] = l.x; cd[0][4] = l.y; cd[0][5] = l.z;
          } else {
            cd[0][0] = T(0); cd[0][1] = T(0); cd[0][2] = T(0); cd[0][3] = ax.x; cd[0][4] = ax.y; cd[0][5] = ax.z;
          }
        }
        for (int k = 0; k < cnt; ++k) {
          T dd[6];
          motion_cross(v, cd[k], dd);
          const T qv = d.get(oQvel + first + k);
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            d.set(oCdof + 6 * (first + k) + c, cd[k][c]);
            if (oCdofdot >= 0) d.set(oCdofdot + 6 * (first + k) + c, dd[c]);
            a[c] += dd[c] * qv;
          }
        }
        for (int k = 0; k < cnt; ++k) {
          const T qv = d.get(oQvel + first + k);
#pragma unroll
          for (int c = 0; c < 6; ++c) v[c] += cd[k][c] * qv;
        }
      }
      T f1[6], f2[6], f3v[6];
      inert_mul(ci, a, f1);
      inert_mul(ci, v, f2);
      motion_cross_force(v, f2, f3v);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        d.set(oCvel + 6 * b + k, v[k]);
        d.set(oCacc + 6 * b + k, a[k]);
        d.set(oCfrc + 6 * b + k, f1[k] + f3v[k]);
        lv[k] = v[k];
        la[k] = a[k];
      }
    }
  }

  // ---- tendon (smooth.py:470-497), fixed tendons only
  if (ntendon) {
    const int oL = m.D(D_ten_length), oJ = m.D(D_ten_J);
    for (int t = 0; t < ntendon; ++t) {
      for (int i = 0; i < nv; ++i) d.set(oJ + t * nv + i, T(0));
      T len = 0;
      for (int w = m.I(I_ten_adr)[t], we = w + m.I(I_ten_num)[t]; w < we; ++w) {
        const int j = m.I(I_wrap_jnt)[w];
        if (j < 0) continue;
        const T c = m.F(F_wrap_prm)[w];
        len += c * d.get(oQpos + jnt_qposadr[j]);
        d.set(oJ + t * nv + jnt_dofadr[j], c);
      }
      d.set(oL + t, len);
    }
  }

  // ---- pass 4, leaves -> root: composite inertia (smooth.py:291-308), the backward half of rne (:453-467) and,
  // while a body's crb / cfrc are in registers, its rows of the dense M (support.make_m :50-69) and qfrc_bias.
  // A child that directly follows its parent hands its sums over in registers; every entry of M is written once.
  if (mpacked && ntendon) {   // tendon armature couples dofs outside the tree pattern: start the packed triangle from zeros
    bool any = false;
    for (int t = 0; t < ntendon; ++t) any |= (m.F(F_ten_armature)[t] != T(0));
    if (any)
      for (int p = 0; p < nv * (nv + 1) / 2; ++p) d.set(oM + p, T(0));
  }
  {
    T hc[10], hf[6];   // crb / cfrc of body b+1 when that body's parent is b
    bool have = false;
    for (int b = nbody - 1; b >= 0; --b) {
      T cb[10], cf[6];
#pragma unroll
      for (int k = 0; k < 10; ++k) cb[k] = d.get(oCrb + 10 * b + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) cf[k] = d.get(oCfrc + 6 * b + k);
      if (have) {
#pragma unroll
        for (int k = 0; k < 10; ++k) cb[k] += hc[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) cf[k] += hf[k];
      }
      if (b == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) d.set(oCrb + k, T(0));   // reference zeroes the world body's crb (smooth.py:306)
#pragma unroll
        for (int k = 0; k < 6; ++k) d.set(oCfrc + k, cf[k]);
        break;
      }
#pragma unroll
      for (int k = 0; k < 10; ++k) d.set(oCrb + 10 * b + k, cb[k]);
#pragma unroll
      for (int k = 0; k < 6; ++k) d.set(oCfrc + 6 * b + k, cf[k]);
      const int p = body_parent[b];
      if (p == b - 1) {
#pragma unroll
        for (int k = 0; k < 10; ++k) hc[k] = cb[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) hf[k] = cf[k];
        have = true;
      } else {
#pragma unroll
        for (int k = 0; k < 10; ++k) d.set(oCrb + 10 * p + k, d.get(oCrb + 10 * p + k) + cb[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) d.set(oCfrc + 6 * p + k, d.get(oCfrc + 6 * p + k) + cf[k]);
        have = false;
      }
      // rows of M for this body's dofs, in groups of up to 3 consecutive dofs that share their ancestor walk
      const int i0 = body_dofadr[b], nd = body_dofnum[b];
      for (int g0 = i0; g0 < i0 + nd; g0 += 3) {
        const int gn = (i0 + nd - g0) < 3 ? (i0 + nd - g0) : 3;
        T cd[3][6], f[3][6];
        for (int k = 0; k < 3; ++k) {
          if (k < gn) {
#pragma unroll
            for (int c = 0; c < 6; ++c) cd[k][c] = d.get(oCdof + 6 * (g0 + k) + c);
            inert_mul(cb, cd[k], f[k]);
            T sb = 0;
#pragma unroll
            for (int c = 0; c < 6; ++c) sb += cd[k][c] * cf[c];
            d.set(oBias + g0 + k, sb);
          }
        }
        // pairs inside the group (dofs of one body form a chain: every earlier dof is an ancestor)
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l <= k; ++l)
            if (k < gn) {
              T sm = 0;
#pragma unroll
              for (int c = 0; c < 6; ++c) sm += f[k][c] * cd[l][c];
              if (l == k) sm += m.F(F_dof_armature)[g0 + k];
              if (mpacked) {
                d.set(oM + (((g0 + k) * (g0 + k + 1)) >> 1) + g0 + l, sm);
              } else {
                d.set(oM + (g0 + k) * nv + g0 + l, sm);
                if (l != k) d.set(oM + (g0 + l) * nv + g0 + k, sm);
              }
            }
        // ancestors below the group, zero-filling the gaps between them
        int prev = g0;
        for (int j = dof_parent[g0]; ; j = dof_parent[j]) {
          const int stop = j < 0 ? -1 : j;
          if (!mpacked)
            for (int z = prev - 1; z > stop; --z)
              for (int k = 0; k < 3; ++k)
                if (k < gn) { d.set(oM + (g0 + k) * nv + z, T(0)); d.set(oM + z * nv + g0 + k, T(0)); }
          if (j < 0) break;
          T cj[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) cj[c] = d.get(oCdof + 6 * j + c);
          for (int k = 0; k < 3; ++k)
            if (k < gn) {
              T sm = 0;
#pragma unroll
              for (int c = 0; c < 6; ++c) sm += f[k][c] * cj[c];
              if (mpacked) {
                d.set(oM + (((g0 + k) * (g0 + k + 1)) >> 1) + j, sm);
              } else {
                d.set(oM + (g0 + k) * nv + j, sm);
                d.set(oM + j * nv + g0 + k, sm);
              }
            }
          prev = j;
        }
      }
    }
  }
  // tendon_armature (smooth.py:500-521): M += J^T diag(armature) J
  if (ntendon) {
    const int oJ = m.D(D_ten_J);
    for (int t = 0; t < ntendon; ++t) {
      const T arm = m.F(F_ten_armature)[t];
      if (arm == T(0)) continue;
      for (int i = 0; i < nv; ++i) {
        const T ji = d.get(oJ + t * nv + i);
        if (ji == T(0)) continue;
        if (mpacked) {
          for (int j = 0; j <= i; ++j) {
            const T add = ji * arm * d.get(oJ + t * nv + j);
            if (add != T(0)) d.set(oM + ((i * (i + 1)) >> 1) + j, d.get(oM + ((i * (i + 1)) >> 1) + j) + add);
          }
        } else {
          for (int j = 0; j < nv; ++j) d.set(oM + i * nv + j, d.get(oM + i * nv + j) + ji * arm * d.get(oJ + t * nv + j));
        }
      }
    }
  }
  if (ntendon) {
    for (int t = 0; t < ntendon; ++t) {
      T v = 0;
      for (int i = 0; i < nv; ++i) v += d.get(m.D(D_ten_J) + t * nv + i) * d.get(oQvel + i);
      d.set(m.D(D_ten_velocity) + t, v);
    }
  }

  // ---- transmission (smooth.py:535-591) and actuator_velocity (forward.py:87-95).  The non-zero entries of an
  // actuator's moment row are a short run (joint transmissions) or a scaled tendon Jacobian row; each row of the
  // dense actuator_moment is written exactly once and never read back.
  const int oAlen = m.D(D_actuator_length), oAmom = m.D(D_actuator_moment), oAvel = m.D(D_actuator_velocity);
  auto transmission = [&](int u, int* first, int* cnt, T* mom, T* len_out, T* vel_out) {
    const int trn = m.I(I_act_trntype)[u], id = m.I(I_act_trnid)[u];
    const T* gear = m.F(F_act_gear) + 6 * u;
    T len = 0, vel = 0;
    *first = 0;
    *cnt = 0;
    if (trn == TRN_TENDON) {
      len = d.get(m.D(D_ten_length) + id) * gear[0];
      *cnt = -1;   // dense row: ten_J[id] * gear[0]
      for (int i = 0; i < nv; ++i) vel += d.get(m.D(D_ten_J) + id * nv + i) * gear[0] * d.get(oQvel + i);
    } else {
      const int jt = jnt_type[id], da = jnt_dofadr[id], qa = jnt_qposadr[id];
      *first = da;
      if (jt == JNT_FREE) {
        *cnt = 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) mom[k] = gear[k];
        if (trn == TRN_JOINTINPARENT) {
          Q4<T> q(d.get(oQpos + qa + 3), d.get(oQpos + qa + 4), d.get(oQpos + qa + 5), d.get(oQpos + qa + 6));
          V3<T> ga = qrot(V3<T>(gear[3], gear[4], gear[5]), qconj(q));
          mom[3] = ga.x; mom[4] = ga.y; mom[5] = ga.z;
        }
        for (int k = 0; k < 6; ++k) vel += mom[k] * d.get(oQvel + da + k);
      } else if (jt == JNT_BALL) {
        *cnt = 3;
        Q4<T> q(d.get(oQpos + qa), d.get(oQpos + qa + 1), d.get(oQpos + qa + 2), d.get(oQpos + qa + 3));
        V3<T> ga(gear[0], gear[1], gear[2]);
        if (trn == TRN_JOINTINPARENT) ga = qrot(ga, qconj(q));
        len = dot(q2axisangle(q), ga);
        mom[0] = ga.x; mom[1] = ga.y; mom[2] = ga.z;
        vel = ga.x * d.get(oQvel + da) + ga.y * d.get(oQvel + da + 1) + ga.z * d.get(oQvel + da + 2);
      } else {
        *cnt = 1;
        len = d.get(oQpos + qa) * gear[0];
        mom[0] = gear[0];
        vel = gear[0] * d.get(oQvel + da);
      }
    }
    *len_out = len;
    *vel_out = vel;
  };
  if (ntendon) {
    for (int t = 0; t < ntendon; ++t) {
      T v = 0;
      for (int i = 0; i < nv; ++i) v += d.get(m.D(D_ten_J) + t * nv + i) * d.get(oQvel + i);
      d.set(m.D(D_ten_velocity) + t, v);
    }
  }

  // ---- passive (passive.py:80-200)
  const int oPassive = m.D(D_qfrc_passive), oGravcomp = m.D(D_qfrc_gravcomp);
  for (int i = 0; i < nv; ++i) { d.set(oPassive + i, T(0)); d.set(oGravcomp + i, T(0)); }
  if (!(flags & (DSBL_SPRING | DSBL_DAMPER))) {
    if (!(flags & DSBL_SPRING)) {
      const T* qs = m.F(F_qpos_spring);
      for (int j = 0; j < njnt; ++j) {
        const int jt = jnt_type[j], qa = jnt_qposadr[j], da = jnt_dofadr[j];
        const T k = m.F(F_jnt_stiffness)[j];
        if (jt == JNT_FREE || jt == JNT_BALL) {
          int qq = qa, dd = da;
          if (jt == JNT_FREE) {
            for (int c = 0; c < 3; ++c) d.set(oPassive + da + c, -k * (d.get(oQpos + qa + c) - qs[qa + c]));
            qq = qa + 3;
            dd = da + 3;
          }
          Q4<T> q(d.get(oQpos + qq), d.get(oQpos + qq + 1), d.get(oQpos + qq + 2), d.get(oQpos + qq + 3));
          Q4<T> s(qs[qq], qs[qq + 1], qs[qq + 2], qs[qq + 3]);
          V3<T> r = q2axisangle(qmul(qconj(s), q));  // quat_sub(q, s)
          d.set(oPassive + dd, -k * r.x); d.set(oPassive + dd + 1, -k * r.y); d.set(oPassive + dd + 2, -k * r.z);
        } else {
          d.set(oPassive + da, -k * (d.get(oQpos + qa) - qs[qa]));
        }
      }
    }
    if (!(flags & DSBL_DAMPER))
      for (int i = 0; i < nv; ++i) d.set(oPassive + i, d.get(oPassive + i) - m.F(F_dof_damping)[i] * d.get(oQvel + i));
    for (int t = 0; t < ntendon; ++t) {
      T frc = 0;
      const T len = d.get(m.D(D_ten_length) + t);
      if (!(flags & DSBL_SPRING)) {
        const T below = m.F(F_ten_lengthspring)[2 * t] - len, above = m.F(F_ten_lengthspring)[2 * t + 1] - len;
        const T ks = m.F(F_ten_stiffness)[t];
        if (below > T(0)) frc = ks * below;
        if (above < T(0)) frc = ks * above;
      }
      if (!(flags & DSBL_DAMPER)) frc += -m.F(F_ten_damping)[t] * d.get(m.D(D_ten_velocity) + t);
      for (int i = 0; i < nv; ++i) d.set(oPassive + i, d.get(oPassive + i) + d.get(m.D(D_ten_J) + t * nv + i) * frc);
    }
    // gravity compensation (passive.py:148-155,187-189)
    if (m.h(H_has_gravcomp) && !(flags & DSBL_GRAVITY)) {
      for (int b = 1; b < nbody; ++b) {
        const T gc = m.F(F_body_gravcomp)[b] * body_mass[b];
        if (gc == T(0)) continue;
        const V3<T> force(-m.opt(OPT_gravity0) * gc, -m.opt(OPT_gravity1) * gc, -m.opt(OPT_gravity2) * gc);
        V3<T> pt, rc;
        ld3(d, oXipos + 3 * b, pt);
        ld3(d, oCom + 3 * body_root[b], rc);
        const V3<T> off = pt - rc;
        for_ancestor_dofs(m, b, [&](int i) {
          V3<T> a, l;
          ld3(d, oCdof + 6 * i, a);
          ld3(d, oCdof + 6 * i + 3, l);
          d.set(oGravcomp + i, d.get(oGravcomp + i) + dot(l + cross(a, off), force));
        });
      }
      for (int i = 0; i < nv; ++i) {
        const T agc = T(m.I(I_jnt_actgravcomp)[m.I(I_dof_jntid)[i]]);
        d.set(oPassive + i, d.get(oPassive + i) + d.get(oGravcomp + i) * (T(1) - agc));
      }
    }
  }

  // ---- fluid forces, inertia-box model (passive.py:30-77,158-173)
  if (m.h(H_has_fluid) && !(flags & (DSBL_SPRING | DSBL_DAMPER))) {
    const T density = m.opt(OPT_density), viscosity = m.opt(OPT_viscosity);
    const V3<T> wind(m.opt(OPT_wind0), m.opt(OPT_wind1), m.opt(OPT_wind2));
    const T pi = T(3.14159265358979323846);
    for (int b = 1; b < nbody; ++b) {
      const T ms = body_mass[b];
      if (!(ms > T(0))) continue;
      const T* I3 = m.F(F_body_inertia) + 3 * b;
      T box[3];
      box[0] = I3[1] + I3[2] - I3[0]; box[1] = I3[0] + I3[2] - I3[1]; box[2] = I3[0] + I3[1] - I3[2];
      for (int k = 0; k < 3; ++k) box[k] = t_sqrt(T(6) * t_max(box[k], T(1e-12)) / t_max(ms, T(1e-12)));
      T R[9], cv[6];
      for (int k = 0; k < 9; ++k) R[k] = d.get(oXimat + 9 * b + k);
      for (int k = 0; k < 6; ++k) cv[k] = d.get(oCvel + 6 * b + k);
      V3<T> ip, rc;
      ld3(d, oXipos + 3 * b, ip);
      ld3(d, oCom + 3 * body_root[b], rc);
      const V3<T> off = ip - rc;
      // transform_motion (math.py:437-452): local = R^T (lin - off x ang), R^T ang
      const V3<T> ang(cv[0], cv[1], cv[2]), lin0 = V3<T>(cv[3], cv[4], cv[5]) - cross(off, ang);
      auto rt = [&](const V3<T>& v) { return V3<T>(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
                                                    R[2] * v.x + R[5] * v.y + R[8] * v.z); };
      auto rr = [&](const V3<T>& v) { return V3<T>(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z,
                                                    R[6] * v.x + R[7] * v.y + R[8] * v.z); };
      const V3<T> la = rt(ang), ll = rt(lin0) - rt(wind);
      const T diam = (box[0] + box[1] + box[2]) / T(3);
      const T lav[3] = {la.x, la.y, la.z}, llv[3] = {ll.x, ll.y, ll.z};
      const T sv[3] = {box[1] * box[2], box[0] * box[2], box[0] * box[1]};
      const T b4[3] = {box[0] * box[0] * box[0] * box[0], box[1] * box[1] * box[1] * box[1], box[2] * box[2] * box[2] * box[2]};
      const T sa[3] = {box[0] * (b4[1] + b4[2]), box[1] * (b4[0] + b4[2]), box[2] * (b4[0] + b4[1])};
      T fa[3], fv[3];
      for (int k = 0; k < 3; ++k) {
        fa[k] = lav[k] * -pi * diam * diam * diam * viscosity - (T(1) * density * sa[k] * t_abs(lav[k]) * lav[k] / T(64));
        fv[k] = llv[k] * T(-3) * pi * diam * viscosity - T(0.5) * density * sv[k] * t_abs(llv[k]) * llv[k];
      }
      const V3<T> force = rr(V3<T>(fv[0], fv[1], fv[2])), torque = rr(V3<T>(fa[0], fa[1], fa[2]));
      for_ancestor_dofs(m, b, [&](int i) {
        V3<T> a, l;
        ld3(d, oCdof + 6 * i, a);
        ld3(d, oCdof + 6 * i + 3, l);
        d.set(oPassive + i, d.get(oPassive + i) + dot(l + cross(a, off), force) + dot(a, torque));
      });
    }
  }

  // ---- actuation (forward.py:102-219)
  const int oAct = m.D(D_act), oActdot = m.D(D_act_dot), oQfa = m.D(D_qfrc_actuator), oForce = m.D(D_actuator_force);
  for (int i = 0; i < nv; ++i) d.set(oQfa + i, T(0));
  for (int i = 0; i < na; ++i) d.set(oActdot + i, T(0));
  const bool actuate = nu && !(flags & DSBL_ACTUATION);
  for (int u = 0; u < nu; ++u) {   // transmission outputs exist even when actuation is disabled
    int mfirst, mcnt;
    T mom[6], len, vel;
    transmission(u, &mfirst, &mcnt, mom, &len, &vel);
    if (oAlen >= 0) d.set(oAlen + u, len);
    if (oAvel >= 0) d.set(oAvel + u, vel);
    if (oAmom < 0) {
      // state layout without a reader of the dense moment rows (ImplicitFast with ball / free / tendon transmissions keeps them)
    } else if (mcnt < 0) {
      const T g0 = m.F(F_act_gear)[6 * u];
      const int id = m.I(I_act_trnid)[u];
      for (int i = 0; i < nv; ++i) d.set(oAmom + u * nv + i, d.get(m.D(D_ten_J) + id * nv + i) * g0);
    } else {
      for (int i = 0; i < nv; ++i) {
        const int k = i - mfirst;
        T v = T(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c == k && c < mcnt) v = mom[c];
        d.set(oAmom + u * nv + i, v);
      }
    }
    if (actuate) {
      T ctrl = in_get<T>(in.ctrl, B, env, u);
      if (!(flags & DSBL_CLAMPCTRL) && m.I(I_act_ctrllimited)[u])
        ctrl = t_min(t_max(ctrl, m.F(F_act_ctrlrange)[2 * u]), m.F(F_act_ctrlrange)[2 * u + 1]);
      T ctrl_act = ctrl;
      const int adr = m.I(I_act_actadr)[u];
      if (adr >= 0) {
        const int dyn = m.I(I_act_dyntype)[u];
        const T* prm = m.F(F_act_dynprm) + 10 * u;
        const T act = d.get(oAct + adr);
        T ad = 0;
        if (dyn == DYN_INTEGRATOR) ad = ctrl;
        else if (dyn == DYN_FILTER || dyn == DYN_FILTEREXACT) ad = (ctrl - act) / t_max(prm[0], T(MJT_MINVAL));
        else if (dyn == DYN_MUSCLE) {
          // support.muscle_dynamics :217-228
          const T cc = t_clamp(ctrl, T(0), T(1)), ac = t_clamp(act, T(0), T(1));
          const T tau_act = prm[0] * (T(0.5) + T(1.5) * ac), tau_deact = prm[1] / (T(0.5) + T(1.5) * ac);
          const T dctrl = cc - act;
          T tau;
          if (prm[2] < T(MJT_MINVAL)) tau = dctrl > T(0) ? tau_act : tau_deact;
          else {
            T x = dctrl / (prm[2] + (prm[2] == T(0) ? T(MJT_MINVAL) : T(0))) + T(0.5);
            T sg = x * x * x * (T(3) * x * (T(2) * x - T(5)) + T(10));
            if (x <= T(0)) sg = T(0);
            if (x >= T(1)) sg = T(1);
            tau = tau_deact + (tau_act - tau_deact) * sg;
          }
          ad = dctrl / t_max(tau, T(MJT_MINVAL));
        }
        d.set(oActdot + adr, ad);
        ctrl_act = d.get(oAct + adr + m.I(I_act_actnum)[u] - 1);
      }
      const int gt = m.I(I_act_gaintype)[u], bt = m.I(I_act_biastype)[u];
      const T* gp = m.F(F_act_gainprm) + 10 * u;
      const T* bp = m.F(F_act_biasprm) + 10 * u;
      T gain = gp[0], bias = 0;
      if (gt == GAIN_AFFINE) gain = gp[0] + gp[1] * len + gp[2] * vel;
      if (bt == BIAS_AFFINE) bias = bp[0] + bp[1] * len + bp[2] * vel;
      if (gt == GAIN_MUSCLE || bt == BIAS_MUSCLE) {
        // support.muscle_gain / muscle_bias :231-296: prm = [range0 range1 force scale lmin lmax vmax fpmax fvmax]
        const T* lr = m.F(F_act_lengthrange) + 2 * u;
        const T acc0 = m.F(F_act_acc0)[u];
        const T mn = T(MJT_MINVAL);
        if (gt == GAIN_MUSCLE) {
          T force0 = gp[2];
          if (force0 < T(0)) force0 = gp[3] / t_max(acc0, mn);
          const T L0 = (lr[1] - lr[0]) / t_max(gp[1] - gp[0], mn);
          const T L = gp[0] + (len - lr[0]) / t_max(L0, mn);
          const T V = vel / t_max(L0 * gp[6], mn);
          const T lmin = gp[4], lmax = gp[5], fvmax = gp[8];
          const T a = T(0.5) * (lmin + T(1)), b = T(0.5) * (T(1) + lmax);
          const T x0 = (L - lmin) / t_max(a - lmin, mn), x1 = (T(1) - L) / t_max(T(1) - a, mn);
          const T x2 = (L - T(1)) / t_max(b - T(1), mn), x3 = (lmax - L) / t_max(lmax - b, mn);
          T FL = L <= b ? T(1) - T(0.5) * x2 * x2 : T(0.5) * x3 * x3;
          if (L <= T(1)) FL = T(1) - T(0.5) * x1 * x1;
          if (L <= a) FL = T(0.5) * x0 * x0;
          if (!(lmin <= L && L <= lmax)) FL = T(0);
          const T y = fvmax - T(1);
          T FV = V <= y ? fvmax - (y - V) * (y - V) / t_max(y, mn) : fvmax;
          if (V <= T(0)) FV = (V + T(1)) * (V + T(1));
          if (V <= T(-1)) FV = T(0);
          gain = -force0 * FL * FV;
        }
        if (bt == BIAS_MUSCLE) {
          T force0 = bp[2];
          if (force0 < T(0)) force0 = bp[3] / t_max(acc0, mn);
          const T L0 = (lr[1] - lr[0]) / t_max(bp[1] - bp[0], mn);
          const T L = bp[0] + (len - lr[0]) / t_max(L0, mn);
          const T lmax = bp[5], fpmax = bp[7];
          const T b = T(0.5) * (T(1) + lmax);
          const T x = (L - T(1)) / t_max(b - T(1), mn);
          bias = L <= b ? -force0 * fpmax * T(0.5) * x * x : -force0 * fpmax * (T(0.5) + (L - b) / t_max(b - T(1), mn));
          if (L <= T(1)) bias = T(0);
        }
      }
      T force = gain * ctrl_act + bias;
      if (m.I(I_act_forcelimited)[u]) force = t_min(t_max(force, m.F(F_act_forcerange)[2 * u]), m.F(F_act_forcerange)[2 * u + 1]);
      if (oForce >= 0) d.set(oForce + u, force);
      if (force != T(0)) {
        if (mcnt < 0) {
          const T g0 = m.F(F_act_gear)[6 * u];
          const int id = m.I(I_act_trnid)[u];
          for (int i = 0; i < nv; ++i) {
            const T mm = d.get(m.D(D_ten_J) + id * nv + i) * g0;
            if (mm != T(0)) d.set(oQfa + i, d.get(oQfa + i) + mm * force);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 6; ++c)
            if (c < mcnt && mom[c] != T(0)) d.set(oQfa + mfirst + c, d.get(oQfa + mfirst + c) + mom[c] * force);
        }
      }
    } else if (oForce >= 0) {
      d.set(oForce + u, T(0));
    }
  }
  if (actuate) {
    for (int i = 0; i < nv; ++i) {
      const int j = m.I(I_dof_jntid)[i];
      T q = d.get(oQfa + i);
      if (m.h(H_has_gravcomp)) q += d.get(oGravcomp + i) * T(m.I(I_jnt_actgravcomp)[j]);
      if (m.I(I_jnt_actfrclimited)[j]) q = t_min(t_max(q, m.F(F_jnt_actfrcrange)[2 * j]), m.F(F_jnt_actfrcrange)[2 * j + 1]);
      d.set(oQfa + i, q);
    }
  }

  // ---- qfrc_smooth (forward.py:222-228; qacc_smooth is solved in k_solve where M is factorised)
  const int oSmooth = m.D(D_qfrc_smooth);
  for (int i = 0; i < nv; ++i)
    d.set(oSmooth + i, d.get(oPassive + i) - d.get(oBias + i) + d.get(oQfa + i) + in_get<T>(in.qfrc_applied, B, env, i));
  for (int b = 1; b < nbody; ++b) {  // support.xfrc_accumulate :184-194 (exact zeros skipped)
    T x[6];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) { x[k] = in_get<T>(in.xfrc_applied, B, env, 6 * b + k); any |= (x[k] != T(0)); }
    if (!any) continue;
    V3<T> pt, rc;
    ld3(d, oXipos + 3 * b, pt);
    ld3(d, oCom + 3 * body_root[b], rc);
    const V3<T> off = pt - rc, force(x[0], x[1], x[2]), torque(x[3], x[4], x[5]);
    for_ancestor_dofs(m, b, [&](int i) {
      V3<T> a, l;
      ld3(d, oCdof + 6 * i, a);
      ld3(d, oCdof + 6 * i + 3, l);
      d.set(oSmooth + i, d.get(oSmooth + i) + dot(l + cross(a, off), force) + dot(a, torque));
    });
  }
}

template <typename T>
__global__ void __launch_bounds__(MJT_ENV_THREADS, MJT_SMOOTH_MINB) k_smooth(const int* __restrict__ gblob, int blob_words, MjtIn in, T* __restrict__ out, int B, int mode) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  EnvBuf<T> d{out + env, (size_t)B};
  smooth_env<T>(m, in, d, (size_t)B, env, mode);
}
