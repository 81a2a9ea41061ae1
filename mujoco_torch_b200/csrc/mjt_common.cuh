// Shared device helpers for the mujoco_torch_b200 kernels (sm_100a).
//
// Layout recap (see layout.py): the ModelBlob is one int32 buffer [header | int arrays | float arrays];
// every CTA stages it into shared memory once with a TMA bulk copy (cp.async.bulk + mbarrier).
// Per-env state and derived quantities live in an env-minor (W, B) buffer: word k of env e is buf[k*B + e],
// so a warp of 32 consecutive envs touches 128 contiguous bytes per access.
#pragma once
// thread-per-env kernels: threads per CTA and minimum resident CTAs per SM (register cap) -- tuning knobs
#ifndef MJT_ENV_THREADS
#define MJT_ENV_THREADS 128
#endif
#ifndef MJT_COLLISION_MINB
#define MJT_COLLISION_MINB 1   // resident CTAs per SM the collision kernel is compiled for (register cap)
#endif
#ifndef MJT_SMOOTH_MINB
#define MJT_SMOOTH_MINB 4
#endif
#ifndef MJT_CONSTRAINT_MINB
#define MJT_CONSTRAINT_MINB 4
#endif
#include <cuda_runtime.h>
#include <stdint.h>
#include "mjt_layout.h"

#define MJT_MINVAL 1e-15
#define MJT_MAXVAL 1e10
#define MJT_MINIMP 1e-4
#define MJT_MAXIMP 0.9999

enum {
  DSBL_CONSTRAINT = 1 << 0, DSBL_EQUALITY = 1 << 1, DSBL_FRICTIONLOSS = 1 << 2, DSBL_LIMIT = 1 << 3, DSBL_CONTACT = 1 << 4,
  DSBL_SPRING = 1 << 5, DSBL_DAMPER = 1 << 6, DSBL_GRAVITY = 1 << 7, DSBL_CLAMPCTRL = 1 << 8, DSBL_WARMSTART = 1 << 9,
  DSBL_FILTERPARENT = 1 << 10, DSBL_ACTUATION = 1 << 11, DSBL_REFSAFE = 1 << 12, DSBL_SENSOR = 1 << 13, DSBL_EULERDAMP = 1 << 15
};
enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_HFIELD = 1, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { INT_EULER = 0, INT_RK4 = 1, INT_IMPLICITFAST = 3 };
enum { SOL_CG = 1, SOL_NEWTON = 2 };
enum { TRN_JOINT = 0, TRN_JOINTINPARENT = 1, TRN_TENDON = 3 };
enum { DYN_NONE = 0, DYN_INTEGRATOR = 1, DYN_FILTER = 2, DYN_FILTEREXACT = 3, DYN_MUSCLE = 4 };
enum { GAIN_FIXED = 0, GAIN_AFFINE = 1, GAIN_MUSCLE = 2 };
enum { BIAS_NONE = 0, BIAS_AFFINE = 1, BIAS_MUSCLE = 2 };

// run-mode flags shared with the host API: the low two bits select step / forward; MJT_MODE_CHECK runs the reference's
// _check_state sanitisation (forward.py:44-59: once, at the start of step(); not in forward() nor in the RK4 sub-stages);
// MJT_MODE_FIXED = step(..., fixed_iterations=True): the solver and its line search run their configured iteration counts
// with no termination test (solver.py:484-487, 536-537)
enum { MJT_MODE_STEP = 1, MJT_MODE_FORWARD = 2, MJT_MODE_KIND = 3, MJT_MODE_CHECK = 4, MJT_MODE_FIXED = 8 };

// ---------------------------------------------------------------------------------------------------------
// Input state: device pointers to env-minor arrays with batch stride B (one (n, B) array per field).
struct MjtIn {
  const void* qpos; const void* qvel; const void* act; const void* time; const void* qacc; const void* qacc_warmstart;
  const void* ctrl; const void* qfrc_applied; const void* xfrc_applied; const void* mocap_pos; const void* mocap_quat;
  const void* sensordata; const void* qfrc_constraint; const int* eq_active;
};

// ---------------------------------------------------------------------------------------------------------
// Model view over the shared-memory copy of the blob.
template <typename T>
struct ModelView {
  const int* w;    // blob words
  const T* f;      // float base
  __device__ __forceinline__ int h(int k) const { return w[k]; }
  __device__ __forceinline__ const int* I(int k) const { return w + w[MJT_OFF_I + k]; }
  __device__ __forceinline__ const T* F(int k) const { return f + w[MJT_OFF_F + k]; }
  __device__ __forceinline__ int D(int k) const { return w[MJT_OFF_D + k]; }
  __device__ __forceinline__ int P(int k) const { return w[MJT_OFF_P + k]; }   // public-buffer offset (layout.py P_FIELDS)
  __device__ __forceinline__ T opt(int k) const { return F(F_opt)[k]; }
};

// Stage the blob into shared memory with one TMA bulk copy (cp.async.bulk.shared::cluster.global) tracked by an
// mbarrier.  ``smem_blob`` must be 16-byte aligned and ``nwords`` a multiple of 4 (the packer guarantees both).
__device__ __forceinline__ void stage_blob(int* smem_blob, uint64_t* bar, const int* __restrict__ gblob, int nwords) {
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(smem_blob);
  const uint32_t bytes = (uint32_t)nwords * 4u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                 "l"(gblob), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
  // all threads wait for phase 0 to complete
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}

template <typename T>
__device__ __forceinline__ ModelView<T> make_view(const int* smem_blob) {
  ModelView<T> m;
  m.w = smem_blob;
  m.f = reinterpret_cast<const T*>(smem_blob + smem_blob[H_float_base_words]);
  return m;
}

// ---------------------------------------------------------------------------------------------------------
// Env accessor over the env-minor (W, B) buffer.
template <typename T>
struct EnvBuf {
  T* __restrict__ p;  // buf + env
  size_t B;
  __device__ __forceinline__ T& at(int word) const { return p[(size_t)word * B]; }
  __device__ __forceinline__ T get(int word) const { return p[(size_t)word * B]; }
  __device__ __forceinline__ void set(int word, T v) const { p[(size_t)word * B] = v; }
};

template <typename T>
__device__ __forceinline__ T in_get(const void* base, size_t B, int env, int k) {
  // optional inputs may be NULL: they read as zero (include/mjt.h)
  return base ? reinterpret_cast<const T*>(base)[(size_t)k * B + env] : T(0);
}

// ---------------------------------------------------------------------------------------------------------
// scalar math in T
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_rsqrt(T x);
template <> __device__ __forceinline__ float t_rsqrt<float>(float x) { return rsqrtf(x); }
template <> __device__ __forceinline__ double t_rsqrt<double>(double x) { return rsqrt(x); }
template <typename T> __device__ __forceinline__ void t_sincos(T x, T* s, T* c);
template <> __device__ __forceinline__ void t_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void t_sincos<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <typename T> __device__ __forceinline__ T t_atan2(T y, T x);
template <> __device__ __forceinline__ float t_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> __device__ __forceinline__ double t_atan2<double>(double y, double x) { return atan2(y, x); }
template <typename T> __device__ __forceinline__ T t_pow(T x, T y);
template <> __device__ __forceinline__ float t_pow<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double t_pow<double>(double x, double y) { return pow(x, y); }
template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_min(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T t_clamp(T x, T lo, T hi) { return t_min(t_max(x, lo), hi); }
template <typename T> __device__ __forceinline__ bool t_finite(T x) { return isfinite(x); }
template <typename T> __device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ float t_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double t_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T>
struct V3 {
  T x, y, z;
  __device__ __forceinline__ V3() {}
  __device__ __forceinline__ V3(T a, T b, T c) : x(a), y(b), z(c) {}
  __device__ __forceinline__ V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
  __device__ __forceinline__ V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
  __device__ __forceinline__ V3 operator*(T s) const { return V3(x * s, y * s, z * s); }
  __device__ __forceinline__ V3 operator-() const { return V3(-x, -y, -z); }
};
template <typename T> __device__ __forceinline__ T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return V3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// reference math.norm / normalize_with_norm (math.py:196-230): exact 0 for the zero vector, x/(n + 1e-6*(n==0))
template <typename T> __device__ __forceinline__ T norm3(const V3<T>& a) { return t_sqrt(dot(a, a)); }
template <typename T> __device__ __forceinline__ V3<T> normalize3(const V3<T>& a, T* n_out = nullptr) {
  T n = norm3(a);
  if (n_out) *n_out = n;
  T d = n + (n == T(0) ? T(1e-6) : T(0));
  return V3<T>(a.x / d, a.y / d, a.z / d);
}

template <typename T>
struct Q4 {
  T w, x, y, z;
  __device__ __forceinline__ Q4() {}
  __device__ __forceinline__ Q4(T a, T b, T c, T d) : w(a), x(b), y(c), z(d) {}
};
template <typename T> __device__ __forceinline__ Q4<T> qmul(const Q4<T>& u, const Q4<T>& v) {
  return Q4<T>(u.w * v.w - u.x * v.x - u.y * v.y - u.z * v.z, u.w * v.x + u.x * v.w + u.y * v.z - u.z * v.y,
               u.w * v.y - u.x * v.z + u.y * v.w + u.z * v.x, u.w * v.z + u.x * v.y - u.y * v.x + u.z * v.w);
}
template <typename T> __device__ __forceinline__ Q4<T> qconj(const Q4<T>& q) { return Q4<T>(q.w, -q.x, -q.y, -q.z); }
template <typename T> __device__ __forceinline__ Q4<T> qnormalize(const Q4<T>& q) {
  T n = t_sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  T d = n + (n == T(0) ? T(1e-6) : T(0));
  return Q4<T>(q.w / d, q.x / d, q.y / d, q.z / d);
}
// reference math.rotate (math.py:246-261)
template <typename T> __device__ __forceinline__ V3<T> qrot(const V3<T>& v, const Q4<T>& q) {
  V3<T> u(q.x, q.y, q.z);
  T s = q.w;
  V3<T> r = u * (T(2) * dot(u, v)) + v * (s * s - dot(u, u));
  return r + cross(u, v) * (T(2) * s);
}
template <typename T> __device__ __forceinline__ void q2mat(const Q4<T>& q, T* m) {
  T ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  m[0] = ww + xx - yy - zz; m[1] = T(2) * (q.x * q.y - q.w * q.z); m[2] = T(2) * (q.x * q.z + q.w * q.y);
  m[3] = T(2) * (q.x * q.y + q.w * q.z); m[4] = ww - xx + yy - zz; m[5] = T(2) * (q.y * q.z - q.w * q.x);
  m[6] = T(2) * (q.x * q.z - q.w * q.y); m[7] = T(2) * (q.y * q.z + q.w * q.x); m[8] = ww - xx - yy + zz;
}
template <typename T> __device__ __forceinline__ Q4<T> axis_angle_q(const V3<T>& axis, T angle) {
  T s, c;
  t_sincos(angle * T(0.5), &s, &c);
  return Q4<T>(c, axis.x * s, axis.y * s, axis.z * s);
}
// reference math.quat_to_axis_angle (math.py:354-360); returns axis*angle
template <typename T> __device__ __forceinline__ V3<T> q2axisangle(const Q4<T>& q, V3<T>* axis_out = nullptr, T* angle_out = nullptr) {
  T sn;
  V3<T> axis = normalize3(V3<T>(q.x, q.y, q.z), &sn);
  T angle = T(2) * t_atan2(sn, q.w);
  const T pi = T(3.14159265358979323846);
  if (angle > pi) angle -= T(2) * pi;
  if (axis_out) *axis_out = axis;
  if (angle_out) *angle_out = angle;
  return axis * angle;
}
// reference math.quat_integrate (math.py:377-383)
template <typename T> __device__ __forceinline__ Q4<T> qintegrate(const Q4<T>& q, const V3<T>& v, T dt) {
  T n;
  V3<T> ax = normalize3(v, &n);
  Q4<T> r = qmul(q, axis_angle_q(ax, dt * n));
  return qnormalize(r);
}
// reference math.orthogonals / make_frame (math.py:485-500); a is assumed already normalised by the caller's normalize3
template <typename T> __device__ __forceinline__ void make_frame(const V3<T>& a_in, T* fr) {
  V3<T> a = normalize3(a_in);
  V3<T> b = (a.y > T(-0.5) && a.y < T(0.5)) ? V3<T>(0, 1, 0) : V3<T>(0, 0, 1);
  b = b - a * dot(a, b);
  b = normalize3(b);
  bool any = (a.x != T(0)) || (a.y != T(0)) || (a.z != T(0));
  if (!any) b = V3<T>(0, 0, 0);
  V3<T> c = cross(a, b);
  fr[0] = a.x; fr[1] = a.y; fr[2] = a.z; fr[3] = b.x; fr[4] = b.y; fr[5] = b.z; fr[6] = c.x; fr[7] = c.y; fr[8] = c.z;
}

// inert_mul (math.py:415-429): i = [Ixx,Iyy,Izz,Ixy,Ixz,Iyz, mx,my,mz, m], v = [ang, lin]
template <typename T> __device__ __forceinline__ void inert_mul(const T* i, const T* v, T* out) {
  V3<T> w(v[0], v[1], v[2]), l(v[3], v[4], v[5]), p(i[6], i[7], i[8]);
  V3<T> c1 = cross(p, l), c2 = cross(p, w);
  out[0] = i[0] * w.x + i[3] * w.y + i[4] * w.z + c1.x;
  out[1] = i[3] * w.x + i[1] * w.y + i[5] * w.z + c1.y;
  out[2] = i[4] * w.x + i[5] * w.y + i[2] * w.z + c1.z;
  out[3] = i[9] * l.x - c2.x;
  out[4] = i[9] * l.y - c2.y;
  out[5] = i[9] * l.z - c2.z;
}
// motion_cross (math.py:455-467)
template <typename T> __device__ __forceinline__ void motion_cross(const T* u, const T* v, T* out) {
  V3<T> ua(u[0], u[1], u[2]), ul(u[3], u[4], u[5]), va(v[0], v[1], v[2]), vl(v[3], v[4], v[5]);
  V3<T> a = cross(ua, va), l = cross(ul, va) + cross(ua, vl);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = l.x; out[4] = l.y; out[5] = l.z;
}
// motion_cross_force (math.py:470-482)
template <typename T> __device__ __forceinline__ void motion_cross_force(const T* v, const T* f, T* out) {
  V3<T> va(v[0], v[1], v[2]), vl(v[3], v[4], v[5]), fa(f[0], f[1], f[2]), fl(f[3], f[4], f[5]);
  V3<T> a = cross(va, fa) + cross(vl, fl), l = cross(va, fl);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = l.x; out[4] = l.y; out[5] = l.z;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
