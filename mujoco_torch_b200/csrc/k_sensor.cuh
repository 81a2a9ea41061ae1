// k_sensor: one thread per env; evaluates the sensors of one stage (pos / vel / acc) and scatters them into
// sensordata (reference sensor.py:56-211, 214-343, 346-440; analytic ray tests ray.py:28-162, 327-372).
// Sensor types the reference has no branch for (touch, ...) keep the caller's value.
// k_rk4_*: small per-env kernels sequencing the RK4 stages (reference forward.py:331-370).
#pragma once
#include "mjt_common.cuh"
#include "k_smooth.cuh"

enum { SENS_FORCE = 4, SENS_TORQUE = 5, SENS_TENDONACTFRC = 17, SENS_SUBTREELINVEL = 36, SENS_SUBTREEANGMOM = 37,
       SENS_ACCELEROMETER = 1, SENS_VELOCIMETER = 2, SENS_GYRO = 3, SENS_MAGNETOMETER = 6, SENS_RANGEFINDER = 7, SENS_JOINTPOS = 9,
       SENS_JOINTVEL = 10, SENS_TENDONPOS = 11, SENS_TENDONVEL = 12, SENS_ACTUATORPOS = 13, SENS_ACTUATORVEL = 14, SENS_ACTUATORFRC = 15,
       SENS_JOINTACTFRC = 16, SENS_BALLQUAT = 18, SENS_BALLANGVEL = 19, SENS_FRAMEPOS = 26, SENS_FRAMEQUAT = 27, SENS_FRAMEXAXIS = 28,
       SENS_FRAMEYAXIS = 29, SENS_FRAMEZAXIS = 30, SENS_FRAMELINVEL = 31, SENS_FRAMEANGVEL = 32, SENS_SUBTREECOM = 35, SENS_CLOCK = 45 };
enum { OBJ_UNKNOWN = 0, OBJ_BODY = 1, OBJ_XBODY = 2, OBJ_GEOM = 5, OBJ_SITE = 6, OBJ_CAMERA = 7 };   // mjtObj

template <typename T> __device__ __forceinline__ T safe_div(T a, T b) { return a / (b + (b == T(0) ? T(MJT_MINVAL) : T(0))); }

template <typename T>
__device__ __forceinline__ void ray_quad(T a, T b, T c, T* x0, T* x1) {
  const T det = b * b - a * c;
  const T det2 = t_sqrt(det);
  T r0 = safe_div(-b - det2, a), r1 = safe_div(-b + det2, a);
  if (det < T(MJT_MINVAL) || r0 < T(0)) r0 = t_inf<T>();
  if (det < T(MJT_MINVAL) || r1 < T(0)) r1 = t_inf<T>();
  *x0 = r0; *x1 = r1;
}

template <typename T>
__device__ T ray_geom(int type, const T* size, const V3<T>& p, const V3<T>& v) {
  const T INF = t_inf<T>();
  if (type == GEOM_PLANE) {
    const T x = -safe_div(p.z, v.z);
    bool valid = v.z <= -T(MJT_MINVAL);
    valid = valid && (x >= T(0));
    const T px = p.x + x * v.x, py = p.y + x * v.y;
    valid = valid && ((size[0] <= T(0)) || (t_abs(px) <= size[0])) && ((size[1] <= T(0)) || (t_abs(py) <= size[1]));
    return valid ? x : INF;
  }
  if (type == GEOM_SPHERE) {
    T x0, x1;
    ray_quad(dot(v, v), dot(v, p), dot(p, p) - size[0] * size[0], &x0, &x1);
    return isinf(x0) ? x1 : x0;
  }
  if (type == GEOM_CAPSULE) {
    T x0, x1;
    ray_quad(v.x * v.x + v.y * v.y, v.x * p.x + v.y * p.y, p.x * p.x + p.y * p.y - size[0] * size[0], &x0, &x1);
    T x = isinf(x0) ? x1 : x0;
    if (!(t_abs(p.z + x * v.z) <= size[1])) x = INF;
    V3<T> dif(p.x, p.y, p.z - size[1]);
    ray_quad(dot(v, v), dot(v, dif), dot(dif, dif) - size[0] * size[0], &x0, &x1);
    if ((p.z + x0 * v.z >= size[1]) && (x0 < x)) x = x0;
    if ((p.z + x1 * v.z >= size[1]) && (x1 < x)) x = x1;
    dif = V3<T>(p.x, p.y, p.z + size[1]);
    ray_quad(dot(v, v), dot(v, dif), dot(dif, dif) - size[0] * size[0], &x0, &x1);
    if ((p.z + x0 * v.z <= -size[1]) && (x0 < x)) x = x0;
    if ((p.z + x1 * v.z <= -size[1]) && (x1 < x)) x = x1;
    return x;
  }
  if (type == GEOM_BOX) {
    const T pp[3] = {p.x, p.y, p.z}, vv[3] = {v.x, v.y, v.z};
    const int ifc[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    T best = INF;
    for (int side = 0; side < 2; ++side)
      for (int a = 0; a < 3; ++a) {
        const T x = side == 0 ? safe_div(size[a] - pp[a], vv[a]) : -safe_div(size[a] + pp[a], vv[a]);
        const int i0 = ifc[a][0], i1 = ifc[a][1];
        const T p0 = pp[i0] + x * vv[i0], p1 = pp[i1] + x * vv[i1];
        const bool valid = (t_abs(p0) <= size[i0]) && (t_abs(p1) <= size[i1]) && (x >= T(0));
        if (valid && x < best) best = x;
      }
    return best;
  }
  return INF;
}

template <typename T> __device__ __forceinline__ T cutoff(T x, T c, int datatype) {
  if (c > T(0)) {
    if (datatype == 0) return t_min(t_max(x, -c), c);
    if (datatype == 1) return t_min(x, c);
  }
  return x;
}

// world position / rotation of a sensor object by mjtObj type (reference sensor.py:63-74)
template <typename T>
__device__ __forceinline__ void sensor_frame(const ModelView<T>& m, const EnvBuf<T>& d, int otype, int id, V3<T>* pos, T* mat) {
  int oP = -1, oR = -1;
  if (otype == OBJ_BODY) { oP = m.D(D_xipos); oR = m.D(D_ximat); }
  else if (otype == OBJ_XBODY) { oP = m.D(D_xpos); oR = m.D(D_xmat); }
  else if (otype == OBJ_GEOM) { oP = m.D(D_geom_xpos); oR = m.D(D_geom_xmat); }
  else if (otype == OBJ_SITE) { oP = m.D(D_site_xpos); oR = m.D(D_site_xmat); }
  if (oP < 0 || id < 0) {
    *pos = V3<T>(0, 0, 0);
    for (int k = 0; k < 9; ++k) mat[k] = (k % 4 == 0) ? T(1) : T(0);
    return;
  }
  ld3(d, oP + 3 * id, *pos);
  for (int k = 0; k < 9; ++k) mat[k] = d.get(oR + 9 * id + k);
}
// body that carries a sensor object
template <typename T>
__device__ __forceinline__ int sensor_body(const ModelView<T>& m, int otype, int id) {
  if (otype == OBJ_BODY || otype == OBJ_XBODY) return id;
  if (otype == OBJ_GEOM) return m.I(I_geom_bodyid)[id];
  if (otype == OBJ_SITE) return m.I(I_site_bodyid)[id];
  return 0;
}
// world orientation of a sensor object as a quaternion (sensor.py:161-176)
template <typename T>
__device__ __forceinline__ Q4<T> sensor_quat(const ModelView<T>& m, const EnvBuf<T>& d, int otype, int id) {
  if (otype == OBJ_UNKNOWN || id < 0) return Q4<T>(1, 0, 0, 0);
  const int b = sensor_body(m, otype, id);
  const Q4<T> bq(d.get(m.D(D_xquat) + 4 * b), d.get(m.D(D_xquat) + 4 * b + 1), d.get(m.D(D_xquat) + 4 * b + 2), d.get(m.D(D_xquat) + 4 * b + 3));
  if (otype == OBJ_XBODY) return bq;
  const T* lq = otype == OBJ_BODY ? m.F(F_body_iquat) + 4 * id : (otype == OBJ_GEOM ? m.F(F_geom_quat) + 4 * id : m.F(F_site_quat) + 4 * id);
  return qmul(bq, f4(lq));
}

template <typename T>
__device__ void sensor_env(const ModelView<T>& m, const MjtIn& in, const EnvBuf<T>& d, const EnvBuf<T>& pub, size_t B, int env, int stage) {
  const int ns = m.h(H_nsensor), oS = m.P(P_sensordata);   // sensordata lives in the caller's buffer (layout.py P table)
  if (stage == 1) {  // first stage seeds the output with the caller's sensordata (untouched slots persist)
    for (int k = 0; k < m.h(H_nsensordata); ++k) pub.set(oS + k, in_get<T>(in.sensordata, B, env, k));
  }
  if (m.h(H_disableflags) & DSBL_SENSOR) return;
  const int* body_root = m.I(I_body_rootid);
  for (int s = 0; s < ns; ++s) {
    if (m.I(I_sens_stage)[s] != stage) continue;
    const int type = m.I(I_sens_type)[s], obj = m.I(I_sens_objid)[s], adr = m.I(I_sens_adr)[s], dtp = m.I(I_sens_datatype)[s];
    const T cut = m.F(F_sens_cutoff)[s];
    T val[4];
    int dim = 0;
    const int otype = m.I(I_sens_objtype)[s], rtype = m.I(I_sens_reftype)[s], rid = m.I(I_sens_refid)[s];
    auto mat_t = [](const T* R, const V3<T>& v) { return V3<T>(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
                                                            R[2] * v.x + R[5] * v.y + R[8] * v.z); };
    if (type == SENS_TENDONPOS) { val[0] = d.get(m.D(D_ten_length) + obj); dim = 1; }
    else if (type == SENS_TENDONVEL) { val[0] = d.get(m.D(D_ten_velocity) + obj); dim = 1; }
    else if (type == SENS_CLOCK) { val[0] = d.get(m.D(D_time)); dim = 1; }
    else if (type == SENS_SUBTREECOM) { for (int k = 0; k < 3; ++k) val[k] = d.get(m.D(D_subtree_com) + 3 * obj + k); dim = 3; }
    else if (type == SENS_BALLANGVEL) { for (int k = 0; k < 3; ++k) val[k] = d.get(m.D(D_qvel) + m.I(I_jnt_dofadr)[obj] + k); dim = 3; }
    else if (type == SENS_BALLQUAT) {
      const int qa = m.I(I_jnt_qposadr)[obj];
      const Q4<T> q = qnormalize(Q4<T>(d.get(m.D(D_qpos) + qa), d.get(m.D(D_qpos) + qa + 1), d.get(m.D(D_qpos) + qa + 2), d.get(m.D(D_qpos) + qa + 3)));
      val[0] = q.w; val[1] = q.x; val[2] = q.y; val[3] = q.z;
      dim = 4;
    } else if (type == SENS_FRAMEQUAT) {
      Q4<T> q = sensor_quat(m, d, otype, obj);
      if (rid != -1) q = qmul(qconj(sensor_quat(m, d, rtype, rid)), q);   // math.quat_inv = conjugate (math.py:264-273)
      val[0] = q.w; val[1] = q.x; val[2] = q.y; val[3] = q.z;
      dim = 4;
    } else if (type == SENS_FRAMEPOS || type == SENS_FRAMEXAXIS || type == SENS_FRAMEYAXIS || type == SENS_FRAMEZAXIS) {
      V3<T> xp, rp;
      T xm[9], rm[9];
      sensor_frame(m, d, otype, obj, &xp, xm);
      V3<T> v = xp;
      if (type != SENS_FRAMEPOS) {
        const int a = type - SENS_FRAMEXAXIS;
        v = V3<T>(xm[a], xm[3 + a], xm[6 + a]);
      }
      if (rid != -1) {
        sensor_frame(m, d, rtype, rid, &rp, rm);
        v = mat_t(rm, type == SENS_FRAMEPOS ? xp - rp : v);
      }
      val[0] = v.x; val[1] = v.y; val[2] = v.z;
      dim = 3;
    } else if (type == SENS_FRAMELINVEL || type == SENS_FRAMEANGVEL) {   // sensor.py:265-325
      V3<T> xp, rp, ang, lin, sc;
      T xm[9], rm[9];
      sensor_frame(m, d, otype, obj, &xp, xm);
      const int ob = sensor_body(m, otype, obj);
      ld3(d, m.D(D_cvel) + 6 * ob, ang);
      ld3(d, m.D(D_cvel) + 6 * ob + 3, lin);
      ld3(d, m.D(D_subtree_com) + 3 * body_root[ob], sc);
      V3<T> v = type == SENS_FRAMEANGVEL ? ang : lin - cross(xp - sc, ang);
      if (rid > -1) {
        V3<T> angr, linr, scr;
        sensor_frame(m, d, rtype, rid, &rp, rm);
        const int rb = sensor_body(m, rtype, rid);
        ld3(d, m.D(D_cvel) + 6 * rb, angr);
        ld3(d, m.D(D_cvel) + 6 * rb + 3, linr);
        ld3(d, m.D(D_subtree_com) + 3 * body_root[rb], scr);
        if (type == SENS_FRAMEANGVEL) v = mat_t(rm, ang - angr);
        else v = mat_t(rm, v - (linr - cross(rp - scr, angr)) + cross(xp - rp, angr));
      }
      val[0] = v.x; val[1] = v.y; val[2] = v.z;
      dim = 3;
    }
    else if (type == SENS_JOINTPOS) { val[0] = d.get(m.D(D_qpos) + m.I(I_jnt_qposadr)[obj]); dim = 1; }
    else if (type == SENS_JOINTVEL) { val[0] = d.get(m.D(D_qvel) + m.I(I_jnt_dofadr)[obj]); dim = 1; }
    else if (type == SENS_ACTUATORPOS) { val[0] = d.get(m.D(D_actuator_length) + obj); dim = 1; }
    else if (type == SENS_ACTUATORVEL) { val[0] = d.get(m.D(D_actuator_velocity) + obj); dim = 1; }
    else if (type == SENS_ACTUATORFRC) { val[0] = d.get(m.D(D_actuator_force) + obj); dim = 1; }
    else if (type == SENS_JOINTACTFRC) { val[0] = d.get(m.D(D_qfrc_actuator) + m.I(I_jnt_dofadr)[obj]); dim = 1; }
    else if (type == SENS_TENDONACTFRC) {   // sensor.py:422-424: total force of the actuators acting on the tendon
      T f = 0;
      for (int u = 0; u < m.h(H_nu); ++u)
        if (m.I(I_act_trntype)[u] == TRN_TENDON && m.I(I_act_trnid)[u] == obj) f += d.get(m.D(D_actuator_force) + u);
      val[0] = f;
      dim = 1;
    } else if (type == SENS_SUBTREELINVEL || type == SENS_SUBTREEANGMOM || type == SENS_FORCE || type == SENS_TORQUE) {
      // The reference reads d.subtree_linvel / subtree_angmom / cfrc_int here (sensor.py:260-265, 394-416), but nothing in it
      // ever computes them (smooth.subtree_vel and smooth.rne_postconstraint do not exist: the hasattr guards at :223-225 and
      // :356-363 skip them), so they hold make_data's zeros and these sensors read zero.  Same output here.
      val[0] = val[1] = val[2] = T(0);
      dim = 3;
    }
    else if (type == SENS_MAGNETOMETER || type == SENS_VELOCIMETER || type == SENS_GYRO || type == SENS_ACCELEROMETER) {
      T R[9];
      for (int k = 0; k < 9; ++k) R[k] = d.get(m.D(D_site_xmat) + 9 * obj + k);
      auto rt = [&](const V3<T>& v) { return V3<T>(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
                                                    R[2] * v.x + R[5] * v.y + R[8] * v.z); };
      V3<T> o;
      if (type == SENS_MAGNETOMETER) {
        o = rt(V3<T>(m.opt(OPT_magnetic0), m.opt(OPT_magnetic1), m.opt(OPT_magnetic2)));
      } else {
        const int b = m.I(I_site_bodyid)[obj];
        V3<T> w, l, sp, sc;
        ld3(d, m.D(D_cvel) + 6 * b, w);
        ld3(d, m.D(D_cvel) + 6 * b + 3, l);
        ld3(d, m.D(D_site_xpos) + 3 * obj, sp);
        ld3(d, m.D(D_subtree_com) + 3 * body_root[b], sc);
        const V3<T> dif = sp - sc;
        if (type == SENS_GYRO) o = rt(w);
        else if (type == SENS_VELOCIMETER) o = rt(l - cross(dif, w));
        else {
          // the reference never fills cacc (no rne_postconstraint): it stays at its make_data zeros, so only the
          // centripetal correction survives (sensor.py:376-390; mujoco >= 3.3.7 gravity term is zero, :36-38)
          const V3<T> ang = rt(w), lin = rt(l - cross(dif, w));
          o = cross(ang, lin);
        }
      }
      val[0] = o.x; val[1] = o.y; val[2] = o.z;
      dim = 3;
    } else if (type == SENS_RANGEFINDER) {
      V3<T> pnt;
      ld3(d, m.D(D_site_xpos) + 3 * obj, pnt);
      const V3<T> vec(d.get(m.D(D_site_xmat) + 9 * obj + 2), d.get(m.D(D_site_xmat) + 9 * obj + 5), d.get(m.D(D_site_xmat) + 9 * obj + 8));
      const int sb = m.I(I_site_bodyid)[obj];
      T best = t_inf<T>();
      for (int k = 0; k < m.h(H_nrfgeom); ++k) {
        int g = m.I(I_rf_geom)[k];
        const bool visible = g >= 0;
        if (!visible) g = -1 - g;
        if (m.I(I_geom_bodyid)[g] == sb || !visible) continue;
        GeomPose<T> gp;
        ld3(d, m.D(D_geom_xpos) + 3 * g, gp.pos);
        for (int q = 0; q < 9; ++q) gp.mat[q] = d.get(m.D(D_geom_xmat) + 9 * g + q);
        const V3<T> dp = pnt - gp.pos;
        const V3<T> lp(gp.mat[0] * dp.x + gp.mat[3] * dp.y + gp.mat[6] * dp.z, gp.mat[1] * dp.x + gp.mat[4] * dp.y + gp.mat[7] * dp.z,
                       gp.mat[2] * dp.x + gp.mat[5] * dp.y + gp.mat[8] * dp.z);
        const V3<T> lv(gp.mat[0] * vec.x + gp.mat[3] * vec.y + gp.mat[6] * vec.z, gp.mat[1] * vec.x + gp.mat[4] * vec.y + gp.mat[7] * vec.z,
                       gp.mat[2] * vec.x + gp.mat[5] * vec.y + gp.mat[8] * vec.z);
        const T x = ray_geom<T>(m.I(I_geom_type)[g], m.F(F_geom_size) + 3 * g, lp, lv);
        if (x < best) best = x;
      }
      val[0] = isinf(best) ? T(-1) : best;
      dim = 1;
    }
    for (int k = 0; k < dim; ++k) pub.set(oS + adr + k, cutoff(val[k], cut, dtp));
  }
}

template <typename T>
__global__ void __launch_bounds__(MJT_ENV_THREADS) k_sensor(const int* __restrict__ gblob, int blob_words, MjtIn in, T* __restrict__ out, T* __restrict__ pubbuf, int B, int stage_mask) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  EnvBuf<T> d{out + env, (size_t)B}, pub{pubbuf + env, (size_t)B};
  for (int st = 1; st <= 3; ++st)
    if (stage_mask & (1 << st)) sensor_env<T>(m, in, d, pub, (size_t)B, env, st);
}

// ---------------------------------------------------------------------------------------------------------
// RK4 sequencing (reference forward._rungekutta4 :331-370).  `t0` is the caller's (public) buffer of the first forward pass (it
// becomes the returned Data), `prev` the public buffer of the previous stage (fields addressed through the P table), `tail` the stage-input / accumulator arrays:
//   tail = [kqpos nq | kqvel nv | kact na | ktime 1 | sum_qvel nv | sum_qacc nv | sum_actdot na]
template <typename T>
__device__ __forceinline__ void integrate_pos(const ModelView<T>& m, const EnvBuf<T>& src, int oSrc, const EnvBuf<T>& dst, int oDst,
                                              const EnvBuf<T>& vel, int oVel, T vscale, T dt) {
  for (int j = 0; j < m.h(H_njnt); ++j) {
    const int jt = m.I(I_jnt_type)[j], qa = m.I(I_jnt_qposadr)[j], da = m.I(I_jnt_dofadr)[j];
    if (jt == JNT_FREE || jt == JNT_BALL) {
      int qq = qa, dd = da;
      if (jt == JNT_FREE) {
        for (int k = 0; k < 3; ++k) dst.set(oDst + qa + k, src.get(oSrc + qa + k) + dt * (vscale * vel.get(oVel + da + k)));
        qq = qa + 3;
        dd = da + 3;
      }
      Q4<T> q(src.get(oSrc + qq), src.get(oSrc + qq + 1), src.get(oSrc + qq + 2), src.get(oSrc + qq + 3));
      q = qintegrate(q, V3<T>(vscale * vel.get(oVel + dd), vscale * vel.get(oVel + dd + 1), vscale * vel.get(oVel + dd + 2)), dt);
      dst.set(oDst + qq, q.w); dst.set(oDst + qq + 1, q.x); dst.set(oDst + qq + 2, q.y); dst.set(oDst + qq + 3, q.z);
    } else {
      dst.set(oDst + qa, src.get(oSrc + qa) + dt * (vscale * vel.get(oVel + da)));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(MJT_ENV_THREADS) k_rk4_stage(const int* __restrict__ gblob, int blob_words, T* __restrict__ t0, T* __restrict__ prev,
                                                   T* __restrict__ tail, int B, int stage) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  const int nq = m.h(H_nq), nv = m.h(H_nv), na = m.h(H_na);
  const EnvBuf<T> d0{t0 + env, (size_t)B}, dp{prev + env, (size_t)B}, tl{tail + env, (size_t)B};
  const int oKq = 0, oKv = nq, oKa = nq + nv, oKt = nq + nv + na, oSv = oKt + 1, oSa = oSv + nv, oSd = oSa + nv;
  const T A[3] = {T(0.5), T(0.5), T(1.0)}, Bw[4] = {T(1.0 / 6.0), T(1.0 / 3.0), T(1.0 / 3.0), T(1.0 / 6.0)};
  const T h = m.opt(OPT_timestep);
  // stage in {0,1,2}: build the inputs of forward pass stage+1; stage 3: final advance
  if (stage == 0) {
    for (int i = 0; i < nv; ++i) { tl.set(oSv + i, Bw[0] * d0.get(m.P(P_qvel) + i)); tl.set(oSa + i, Bw[0] * d0.get(m.P(P_qacc) + i)); }
    for (int i = 0; i < na; ++i) tl.set(oSd + i, Bw[0] * d0.get(m.P(P_act_dot) + i));
  } else {
    // accumulate the stage that just finished: kqvel (its input velocity, still in tail), its qacc and act_dot
    for (int i = 0; i < nv; ++i) {
      tl.set(oSv + i, tl.get(oSv + i) + Bw[stage] * tl.get(oKv + i));
      tl.set(oSa + i, tl.get(oSa + i) + Bw[stage] * dp.get(m.P(P_qacc) + i));
    }
    for (int i = 0; i < na; ++i) tl.set(oSd + i, tl.get(oSd + i) + Bw[stage] * dp.get(m.P(P_act_dot) + i));
  }
  if (stage < 3) {
    const T a = A[stage];
    // kqpos = integrate(qpos_t0, a * kqvel_prev); kqvel_prev is t0.qvel for stage 0, else tail.kqvel
    if (stage == 0) integrate_pos(m, d0, m.P(P_qpos), tl, oKq, d0, m.P(P_qvel), a, h);
    else integrate_pos(m, d0, m.P(P_qpos), tl, oKq, tl, oKv, a, h);
    for (int i = 0; i < na; ++i) tl.set(oKa + i, d0.get(m.P(P_act) + i) + a * dp.get(m.P(P_act_dot) + i) * h);
    for (int i = 0; i < nv; ++i) tl.set(oKv + i, d0.get(m.P(P_qvel) + i) + a * dp.get(m.P(P_qacc) + i) * h);
    const T C[3] = {T(0.5), T(0.5), T(1.0)};
    tl.set(oKt, d0.get(m.P(P_time)) + C[stage] * h);
  } else {
    // _advance(d_t0, act_dot_sum, qacc_sum, qvel_sum)
    for (int u = 0; u < m.h(H_nu); ++u) {
      const int adr = m.I(I_act_actadr)[u];
      if (adr < 0) continue;
      for (int k = adr; k < adr + m.I(I_act_actnum)[u]; ++k) {
        const T act = d0.get(m.P(P_act) + k), ad = tl.get(oSd + k);
        T nx;
        if (m.I(I_act_dyntype)[u] == DYN_FILTEREXACT) {
          const T tau = t_max(m.F(F_act_dynprm)[10 * u], T(MJT_MINVAL));
          nx = act + ad * tau * (T(1) - t_exp(-h / tau));
        } else nx = act + ad * h;
        if (m.I(I_act_actlimited)[u]) nx = t_clamp(nx, m.F(F_act_actrange)[2 * u], m.F(F_act_actrange)[2 * u + 1]);
        d0.set(m.P(P_act) + k, nx);
      }
    }
    integrate_pos(m, d0, m.P(P_qpos), d0, m.P(P_qpos), tl, oSv, T(1), h);
    for (int i = 0; i < nv; ++i) d0.set(m.P(P_qvel) + i, d0.get(m.P(P_qvel) + i) + tl.get(oSa + i) * h);
    d0.set(m.P(P_time), d0.get(m.P(P_time)) + h);
  }
}
