// k_solve: one WARP per environment; E environments (warps) per CTA share one staged ModelBlob.
// Fuses the reference's factor_m + qacc_smooth solve (smooth.py:311-354, forward.py:222-228), the primal
// CG / Newton solver with exact line search (solver.py:244-553), the Euler-with-damping / ImplicitFast
// re-factorisations (forward.py:313-328, 404-416; derivative.py:22-68) and _advance (forward.py:255-310).
//
// Mapping: lane i owns dof i (and i+32 when nv > 32): every nv-vector (qacc, Ma, grad, search ...) lives in
// registers, one element per lane; dot products and costs are warp-shuffle reductions.  The per-env matrices
// (J, M, and one factor workspace) live in shared memory with odd leading dimensions (bank-conflict-free row
// and column walks).  Because a warp == an environment, every data-dependent branch (active rows, line-search
// bracketing, early exit) is warp-uniform: inactive constraint rows and converged envs cost nothing, which is
// exactly the per-env early exit the reference's masked while_loop emulates.
#pragma once
#include "mjt_common.cuh"
#include "k_smooth.cuh"

struct SolveLayout {
  int ldj, ldm, oJ, oM, oL, oXs, oD, oAref, oFloss, oJaref, oJv, oForce, oAct, words;
};

__host__ __device__ inline SolveLayout solve_layout(int nv, int nefc) {
  SolveLayout s;
  s.ldj = nv | 1;
  s.ldm = nv | 1;
  int o = 0;
  s.oJ = o; o += nefc * s.ldj;
  s.oM = o; o += nv * s.ldm;
  s.oL = o; o += nv * s.ldm;
  s.oXs = o; o += nv + 1;
  s.oD = o; o += nefc;
  s.oAref = o; o += nefc;
  s.oFloss = o; o += nefc;
  s.oJaref = o; o += nefc;
  s.oJv = o; o += nefc;
  s.oForce = o; o += nefc;
  s.oAct = o; o += nefc;  // active flags / compacted active indices (stored as T-sized ints)
  s.words = (o + 3) & ~3;
  return s;
}

template <typename T, int DPL>
struct Vec {
  T v[DPL];
};

// DPL-indexed shuffle helper needs a compile-time register index; specialise the access.
template <typename T>
__device__ __forceinline__ T vec_bcast(const Vec<T, 1>& x, int i) { return __shfl_sync(0xffffffffu, x.v[0], i & 31); }
template <typename T>
__device__ __forceinline__ T vec_bcast(const Vec<T, 2>& x, int i) {
  const T a = __shfl_sync(0xffffffffu, x.v[0], i & 31), b = __shfl_sync(0xffffffffu, x.v[1], i & 31);
  return (i >> 5) ? b : a;
}

template <typename T, int DPL>
struct SolveCtx {
  const ModelView<T>& m;
  T* s;  // this env's shared-memory region
  SolveLayout L;
  int nv, nefc, lane;

  __device__ __forceinline__ T* J() const { return s + L.oJ; }
  __device__ __forceinline__ T* Mx() const { return s + L.oM; }
  __device__ __forceinline__ T* Lx() const { return s + L.oL; }
  __device__ __forceinline__ T* xs() const { return s + L.oXs; }

  __device__ __forceinline__ Vec<T, DPL> zero() const {
    Vec<T, DPL> r;
#pragma unroll
    for (int k = 0; k < DPL; ++k) r.v[k] = T(0);
    return r;
  }
  __device__ __forceinline__ T dotv(const Vec<T, DPL>& a, const Vec<T, DPL>& b) const {
    T p = 0;
#pragma unroll
    for (int k = 0; k < DPL; ++k) p += a.v[k] * b.v[k];
    return warp_sum(p);
  }
  __device__ __forceinline__ void to_smem(const Vec<T, DPL>& a, T* dst) const {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) dst[i] = a.v[k];
    }
    __syncwarp();
  }
  // y = A x for a dense (nv x nv) shared-memory matrix
  __device__ __forceinline__ Vec<T, DPL> matvec(const T* A, const Vec<T, DPL>& x) const {
    to_smem(x, xs());
    Vec<T, DPL> y = zero();
    for (int c = 0; c < nv; ++c) {
      const T xc = xs()[c];
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) y.v[k] += A[i * L.ldm + c] * xc;
      }
    }
    return y;
  }
  // out[r] = J[r,:] . x  (+ optional subtraction of sub[r])
  __device__ __forceinline__ void jmul(const Vec<T, DPL>& x, T* out, const T* sub) const {
    to_smem(x, xs());
    for (int r = lane; r < nefc; r += 32) {
      const T* row = J() + r * L.ldj;
      T acc = 0;
      for (int c = 0; c < nv; ++c) acc += row[c] * xs()[c];
      out[r] = sub ? acc - sub[r] : acc;
    }
    __syncwarp();
  }
  // y = J^T f  (f in shared memory; zero entries skipped, warp-uniformly)
  __device__ __forceinline__ Vec<T, DPL> jtmul(const T* f) const {
    Vec<T, DPL> y = zero();
    for (int r = 0; r < nefc; ++r) {
      const T fr = f[r];
      if (fr == T(0)) continue;
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) y.v[k] += J()[r * L.ldj + i] * fr;
      }
    }
    return y;
  }
  // In-place dense Cholesky of the lower triangle of A (n x n, ld = ldm); upper triangle zeroed.
  // Numeric conventions of reference math.small_cholesky :87-129: n <= 16 clamps each pivot at 1e-12,
  // n > 16 adds 1e-10 to the diagonal (the torch.linalg.cholesky fallback path).
  __device__ __forceinline__ void cholesky(T* A) const {
    const int n = nv, ld = L.ldm;
    const bool big = n > 16;
    for (int j = 0; j < n; ++j) {
      T p = 0;
      for (int k = lane; k < j; k += 32) { const T l = A[j * ld + k]; p += l * l; }
      T sdiag = A[j * ld + j] + (big ? T(1e-10) : T(0)) - warp_sum(p);
      if (!big) sdiag = t_max(sdiag, T(1e-12));
      const T ljj = t_sqrt(sdiag);
      for (int i = j + 1 + lane; i < n; i += 32) {
        T t = A[i * ld + j];
        for (int k = 0; k < j; ++k) t -= A[i * ld + k] * A[j * ld + k];
        A[i * ld + j] = t / ljj;
        A[j * ld + i] = T(0);
      }
      if (lane == 0) A[j * ld + j] = ljj;
      __syncwarp();
    }
  }
  // x <- (L L^T)^-1 x with x distributed one element per lane (reference small_cholesky_solve :132-168)
  __device__ __forceinline__ Vec<T, DPL> cho_solve(const T* Lm, Vec<T, DPL> x) const {
    const int n = nv, ld = L.ldm;
    for (int i = 0; i < n; ++i) {
      const T xi = vec_bcast(x, i) / Lm[i * ld + i];
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int j = lane + 32 * k;
        if (j == i) x.v[k] = xi;
        else if (j > i && j < n) x.v[k] -= Lm[j * ld + i] * xi;
      }
    }
    for (int i = n - 1; i >= 0; --i) {
      const T xi = vec_bcast(x, i) / Lm[i * ld + i];
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int j = lane + 32 * k;
        if (j == i) x.v[k] = xi;
        else if (j < i) x.v[k] -= Lm[i * ld + j] * xi;
      }
    }
    return x;
  }
};

template <typename T, int DPL>
__device__ void solve_env(const ModelView<T>& m, const MjtIn& in, const EnvBuf<T>& d, size_t B, int env, T* sm, int mode) {
  typedef Vec<T, DPL> V;
  const int lane = threadIdx.x & 31;
  const int nv = m.h(H_nv), nefc = m.h(H_nefc), flags = m.h(H_disableflags);
  SolveCtx<T, DPL> c{m, sm, solve_layout(nv, nefc), nv, nefc, lane};
  const SolveLayout& Lo = c.L;
  T* const Jm = c.J();
  T* const Mm = c.Mx();
  T* const Lm = c.Lx();
  T* const D = sm + Lo.oD;
  T* const aref = sm + Lo.oAref;
  T* const floss = sm + Lo.oFloss;
  T* const Jaref = sm + Lo.oJaref;
  T* const jv = sm + Lo.oJv;
  T* const force = sm + Lo.oForce;
  T* const actf = sm + Lo.oAct;
  const int ld = Lo.ldm;

  auto loadv = [&](int off) {
    V r;
#pragma unroll
    for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; r.v[k] = i < nv ? d.get(off + i) : T(0); }
    return r;
  };
  auto storev = [&](int off, const V& x) {
#pragma unroll
    for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; if (i < nv) d.set(off + i, x.v[k]); }
  };
  auto axpy = [&](const V& a, T s, const V& b) {  // a + s*b
    V r;
#pragma unroll
    for (int k = 0; k < DPL; ++k) r.v[k] = a.v[k] + s * b.v[k];
    return r;
  };
  auto neg = [&](const V& a) {
    V r;
#pragma unroll
    for (int k = 0; k < DPL; ++k) r.v[k] = -a.v[k];
    return r;
  };

  // ---- factor_m + qacc_smooth
  for (int idx = lane; idx < nv * nv; idx += 32) {
    const int i = idx / nv, j = idx - i * nv;
    Lm[i * ld + j] = Mm[i * ld + j];
  }
  __syncwarp();
  c.cholesky(Lm);
  for (int idx = lane; idx < nv * nv; idx += 32) {
    const int i = idx / nv, j = idx - i * nv;
    d.set(m.D(D_qLD) + idx, Lm[i * ld + j]);
  }
  const V qfrc_smooth = loadv(m.D(D_qfrc_smooth));
  const V qacc_smooth = c.cho_solve(Lm, qfrc_smooth);
  storev(m.D(D_qacc_smooth), qacc_smooth);

  V qacc = qacc_smooth;
  V qfrc_con = c.zero();

  if (nefc == 0) {
    // forward.py:390-396; qfrc_constraint keeps the caller's value (solve() never runs)
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) {
        qfrc_con.v[k] = in_get<T>(in.qfrc_constraint, B, env, i);
        d.set(m.D(D_qacc_warmstart) + i, in_get<T>(in.qacc_warmstart, B, env, i));
      }
    }
  } else {
    const int ne_nf = m.h(H_ne) + m.h(H_nf);
    const bool use_floss = m.h(H_nf) > 0 && !(flags & DSBL_FRICTIONLOSS);
    const bool newton = m.h(H_solver) == SOL_NEWTON;
    const T scale = m.opt(OPT_meaninertia) * T(nv > 1 ? nv : 1);
    const T tolerance = m.opt(OPT_tolerance), ls_tolerance = m.opt(OPT_ls_tolerance);
    const int iterations = m.h(H_iterations), ls_iterations = m.h(H_ls_iterations);

    V Ma, grad, Mgrad, search;
    T gauss = 0, cost = t_inf<T>(), prev_cost = 0;

    // reference _update_constraint :320-357
    auto update_constraint = [&]() {
      T cpart = 0;
      for (int r = lane; r < nefc; r += 32) {
        const T ja = Jaref[r], dr = D[r];
        bool act = (ja < T(0)) || (r < ne_nf);
        T ff = 0;
        if (use_floss) {
          const T fl = floss[r];
          if (fl > T(0)) {
            const T rr = T(1) / (dr + (dr == T(0) ? T(MJT_MINVAL) : T(0)));
            const bool lin_neg = ja <= -rr * fl, lin_pos = ja >= rr * fl;
            if (lin_neg) { act = false; ff = fl; cpart += -T(0.5) * rr * fl * fl - fl * ja; }
            if (lin_pos) { act = false; ff = -fl; cpart += -T(0.5) * rr * fl * fl + fl * ja; }
          }
        }
        force[r] = (act ? dr * -ja : T(0)) + ff;
        actf[r] = act ? T(1) : T(0);
        if (act) cpart += T(0.5) * dr * ja * ja;
      }
      __syncwarp();
      qfrc_con = c.jtmul(force);
      T g = 0;
#pragma unroll
      for (int k = 0; k < DPL; ++k) g += (Ma.v[k] - qfrc_smooth.v[k]) * (qacc.v[k] - qacc_smooth.v[k]);
      gauss = T(0.5) * warp_sum(g);
      prev_cost = cost;
      cost = warp_sum(cpart) + gauss;
    };

    // reference _update_gradient :359-376
    auto update_gradient = [&]() {
#pragma unroll
      for (int k = 0; k < DPL; ++k) grad.v[k] = Ma.v[k] - qfrc_smooth.v[k] - qfrc_con.v[k];
      if (!newton) {
        Mgrad = c.cho_solve(Lm, grad);  // Lm still holds chol(M)
        return;
      }
      // H = M + J^T diag(D * active) J, accumulated per lane-owned row in registers, structural zeros skipped
      for (int col = 0; col < nv; ++col) {
        T acc[DPL];
#pragma unroll
        for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; acc[k] = i < nv ? Mm[i * ld + col] : T(0); }
        for (int r = 0; r < nefc; ++r) {
          if (actf[r] == T(0)) continue;
          const T jrc = Jm[r * Lo.ldj + col];
          if (jrc == T(0)) continue;
          const T w = D[r] * jrc;
#pragma unroll
          for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; if (i < nv) acc[k] += Jm[r * Lo.ldj + i] * w; }
        }
#pragma unroll
        for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; if (i < nv) Lm[i * ld + col] = acc[k]; }
      }
      __syncwarp();
      c.cholesky(Lm);
      Mgrad = c.cho_solve(Lm, grad);
    };

    // reference _create_context :293-318
    auto create_context = [&](const V& q, bool grad_flag) {
      qacc = q;
      c.jmul(qacc, Jaref, aref);
      Ma = c.matvec(Mm, qacc);
      cost = t_inf<T>();
      prev_cost = 0;
      update_constraint();
      if (grad_flag) {
        update_gradient();
        search = neg(Mgrad);
      }
    };

    // reference _linesearch :378-497
    auto linesearch = [&]() {
      const T smag = t_sqrt(c.dotv(search, search)) * scale;
      const T gtol = tolerance * ls_tolerance * smag;
      const V mv = c.matvec(Mm, search);
      c.jmul(search, jv, nullptr);
      T qg0 = gauss, qg1, qg2;
      {
        T a = 0, b = 0, e = 0;
#pragma unroll
        for (int k = 0; k < DPL; ++k) { a += search.v[k] * Ma.v[k]; b += search.v[k] * qfrc_smooth.v[k]; e += search.v[k] * mv.v[k]; }
        qg1 = warp_sum(a) - warp_sum(b);
        qg2 = T(0.5) * warp_sum(e);
      }
      struct Pt { T alpha, cost, d0, d1; };
      // evaluates up to three step sizes in one pass over the constraint rows
      auto points = [&](const T* al, int n, Pt* out) {
        T q[3][3];
        for (int a = 0; a < 3; ++a) q[a][0] = q[a][1] = q[a][2] = T(0);
        for (int r = lane; r < nefc; r += 32) {
          const T ja = Jaref[r], jr = jv[r], dr = D[r];
          const T c0 = T(0.5) * ja * ja * dr, c1 = jr * ja * dr, c2 = T(0.5) * jr * jr * dr;
          T fl = 0, rf = 0;
          const bool hasf = use_floss && (fl = floss[r]) > T(0);
          if (hasf) rf = fl / (dr + (dr == T(0) ? T(MJT_MINVAL) : T(0)));
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (a >= n) break;
            const T x = ja + al[a] * jr;
            bool act = (x < T(0)) || (r < ne_nf);
            if (hasf) {
              if (x <= -rf) { act = false; q[a][0] += fl * (-T(0.5) * rf - ja); q[a][1] += -fl * jr; }
              if (x >= rf) { act = false; q[a][0] += fl * (-T(0.5) * rf + ja); q[a][1] += fl * jr; }
            }
            if (act) { q[a][0] += c0; q[a][1] += c1; q[a][2] += c2; }
          }
        }
        for (int a = 0; a < n; ++a) {
          const T t0 = qg0 + warp_sum(q[a][0]), t1 = qg1 + warp_sum(q[a][1]), t2 = qg2 + warp_sum(q[a][2]);
          out[a].alpha = al[a];
          out[a].cost = al[a] * al[a] * t2 + al[a] * t1 + t0;
          out[a].d0 = T(2) * al[a] * t2 + t1;
          out[a].d1 = T(2) * t2 + (t2 == T(0) ? T(MJT_MINVAL) : T(0));
        }
      };
      Pt p0, p1, lo, hi;
      T al[3] = {T(0), T(0), T(0)};
      points(al, 1, &p0);
      al[0] = p0.alpha - p0.d0 / p0.d1;
      points(al, 1, &p1);
      const bool early = t_abs(p1.d0) < gtol;
      if (p1.d0 < p0.d0) { hi = p0; lo = p1; } else { hi = p1; lo = p0; }
      bool swap = !early;
      int it = 0;
      while (true) {
        bool done = it >= ls_iterations;
        done = done || !swap;
        done = done || (lo.d0 < T(0) && lo.d0 > -gtol);
        done = done || (hi.d0 > T(0) && hi.d0 < gtol);
        if (done) break;
        Pt cand[3];
        al[0] = lo.alpha - lo.d0 / lo.d1;
        al[1] = hi.alpha - hi.d0 / hi.d1;
        al[2] = T(0.5) * (lo.alpha + hi.alpha);
        points(al, 3, cand);
        const Pt &lo_next = cand[0], &hi_next = cand[1], &mid = cand[2];
        const bool not_bracketed = (lo.d0 < T(0)) == (hi.d0 < T(0));
        auto sw = [&](T cur, T cd) {
          const bool in_bracket = ((cur < cd) && (cd < T(0))) || ((cur > cd) && (cd > T(0)));
          return in_bracket || (not_bracketed && (t_abs(cd) < t_abs(cur)));
        };
        bool s1 = sw(lo.d0, lo_next.d0); if (s1) lo = lo_next;
        bool s2 = sw(lo.d0, mid.d0); if (s2) lo = mid;
        bool s3 = sw(lo.d0, hi_next.d0); if (s3) lo = hi_next;
        bool s4 = sw(hi.d0, hi_next.d0); if (s4) hi = hi_next;
        bool s5 = sw(hi.d0, mid.d0); if (s5) hi = mid;
        bool s6 = sw(hi.d0, lo_next.d0); if (s6) hi = lo_next;
        swap = s1 || s2 || s3 || s4 || s5 || s6;
        ++it;
      }
      const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
      const T alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
      if (improved) {
        qacc = axpy(qacc, alpha, search);
        Ma = axpy(Ma, alpha, mv);
        for (int r = lane; r < nefc; r += 32) Jaref[r] += jv[r] * alpha;
        __syncwarp();
      }
    };

    auto body = [&]() {
      linesearch();
      const V prev_grad = grad, prev_Mgrad = Mgrad;
      update_constraint();
      update_gradient();
      if (newton) {
        search = neg(Mgrad);
      } else {  // Polak-Ribiere :519-523
        V dm;
#pragma unroll
        for (int k = 0; k < DPL; ++k) dm.v[k] = Mgrad.v[k] - prev_Mgrad.v[k];
        T beta = c.dotv(grad, dm) / t_max(c.dotv(prev_grad, prev_Mgrad), T(MJT_MINVAL));
        beta = t_max(beta, T(0));
#pragma unroll
        for (int k = 0; k < DPL; ++k) search.v[k] = -Mgrad.v[k] + beta * search.v[k];
      }
    };

    // warm start :526-531
    V q0 = qacc_smooth;
    if (!(flags & DSBL_WARMSTART)) {
      V warm;
#pragma unroll
      for (int k = 0; k < DPL; ++k) { const int i = lane + 32 * k; warm.v[k] = i < nv ? in_get<T>(in.qacc_warmstart, B, env, i) : T(0); }
      create_context(warm, false);
      const T cw = cost;
      create_context(qacc_smooth, false);
      if (cw < cost) q0 = warm;
    }
    create_context(q0, true);
    if (iterations == 1) {
      body();
    } else {
      int niter = 0;
      while (true) {
        const T improvement = (prev_cost - cost) / scale;
        const T gradient = t_sqrt(c.dotv(grad, grad)) / scale;
        bool done = niter >= iterations;
        done = done || (improvement < tolerance);
        done = done || (gradient < tolerance);
        if (done) break;
        body();
        ++niter;
      }
    }
    storev(m.D(D_qacc_warmstart), qacc);
    for (int r = lane; r < nefc; r += 32) d.set(m.D(D_efc_force) + r, force[r]);
  }
  storev(m.D(D_qacc), qacc);
  storev(m.D(D_qfrc_constraint), qfrc_con);

  if (mode != MJT_MODE_STEP) return;
  const int integrator = m.h(H_integrator);
  if (integrator == INT_RK4) return;  // RK4 is sequenced on the host from forward passes

  // ---- integrator: velocity-implicit re-solve (forward.py:313-328 / 404-416)
  const T dt = m.opt(OPT_timestep);
  V qacc_int = qacc;
  bool refactor = false;
  if (integrator == INT_EULER && !(flags & DSBL_EULERDAMP)) {
    for (int idx = lane; idx < nv * nv; idx += 32) {
      const int i = idx / nv, j = idx - i * nv;
      Lm[i * ld + j] = Mm[i * ld + j] + (i == j ? dt * m.F(F_dof_damping)[i] : T(0));
    }
    refactor = true;
  } else if (integrator == INT_IMPLICITFAST) {
    // qderiv = moment^T diag(biasprm2*[affine] + gainprm2*[affine]*ctrl) moment - diag(damping)  (derivative.py:22-68)
    const int nu = m.h(H_nu);
    const bool actuation = !(flags & DSBL_ACTUATION), damper = !(flags & DSBL_DAMPER);
    if (actuation || damper || m.h(H_ntendon)) {
      for (int idx = lane; idx < nv * nv; idx += 32) {
        const int i = idx / nv, j = idx - i * nv;
        T qd = (damper && i == j) ? -m.F(F_dof_damping)[i] : T(0);
        if (actuation)
          for (int u = 0; u < nu; ++u) {
            T vel = 0;
            if (m.I(I_act_biastype)[u] == BIAS_AFFINE) vel += m.F(F_act_biasprm)[10 * u + 2];
            if (m.I(I_act_gaintype)[u] == GAIN_AFFINE) {
              const T cu = m.I(I_act_dyntype)[u] != DYN_NONE ? d.get(m.D(D_act) + m.I(I_act_actadr)[u]) : in_get<T>(in.ctrl, B, env, u);
              vel += m.F(F_act_gainprm)[10 * u + 2] * cu;
            }
            if (vel == T(0)) continue;
            qd += d.get(m.D(D_actuator_moment) + u * nv + i) * vel * d.get(m.D(D_actuator_moment) + u * nv + j);
          }
        for (int t = 0; t < m.h(H_ntendon); ++t)
          qd -= d.get(m.D(D_ten_J) + t * nv + i) * m.F(F_ten_damping)[t] * d.get(m.D(D_ten_J) + t * nv + j);
        Lm[i * ld + j] = Mm[i * ld + j] - dt * qd;
      }
      refactor = true;
    }
  }
  if (refactor) {
    __syncwarp();
    c.cholesky(Lm);
    V rhs;
#pragma unroll
    for (int k = 0; k < DPL; ++k) rhs.v[k] = qfrc_smooth.v[k] + qfrc_con.v[k];
    qacc_int = c.cho_solve(Lm, rhs);
  }

  // ---- _advance (forward.py:255-310)
  const int na = m.h(H_na), nu = m.h(H_nu);
  for (int u = lane; u < nu; u += 32) {
    const int adr = m.I(I_act_actadr)[u];
    if (adr < 0) continue;
    for (int a = adr; a < adr + m.I(I_act_actnum)[u]; ++a) {
      const T act = d.get(m.D(D_act) + a), ad = d.get(m.D(D_act_dot) + a);
      T nx;
      if (m.I(I_act_dyntype)[u] == DYN_FILTEREXACT) {
        const T tau = t_max(m.F(F_act_dynprm)[10 * u], T(MJT_MINVAL));
        nx = act + ad * tau * (T(1) - t_exp(-dt / tau));
      } else {
        nx = act + ad * dt;
      }
      if (m.I(I_act_actlimited)[u]) nx = t_clamp(nx, m.F(F_act_actrange)[2 * u], m.F(F_act_actrange)[2 * u + 1]);
      d.set(m.D(D_act) + a, nx);
    }
  }
  (void)na;
  V qvel = loadv(m.D(D_qvel));
  qvel = axpy(qvel, dt, qacc_int);
  storev(m.D(D_qvel), qvel);
  c.to_smem(qvel, c.xs());
  const T* nvel = c.xs();
  const int oQ = m.D(D_qpos);
  for (int j = lane; j < m.h(H_njnt); j += 32) {
    const int jt = m.I(I_jnt_type)[j], qa = m.I(I_jnt_qposadr)[j], da = m.I(I_jnt_dofadr)[j];
    if (jt == JNT_FREE || jt == JNT_BALL) {
      int qq = qa, dd = da;
      if (jt == JNT_FREE) {
        for (int k = 0; k < 3; ++k) d.set(oQ + qa + k, d.get(oQ + qa + k) + dt * nvel[da + k]);
        qq = qa + 3;
        dd = da + 3;
      }
      Q4<T> q(d.get(oQ + qq), d.get(oQ + qq + 1), d.get(oQ + qq + 2), d.get(oQ + qq + 3));
      q = qintegrate(q, V3<T>(nvel[dd], nvel[dd + 1], nvel[dd + 2]), dt);
      d.set(oQ + qq, q.w); d.set(oQ + qq + 1, q.x); d.set(oQ + qq + 2, q.y); d.set(oQ + qq + 3, q.z);
    } else {
      d.set(oQ + qa, d.get(oQ + qa) + dt * nvel[da]);
    }
  }
  if (lane == 0) d.set(m.D(D_time), d.get(m.D(D_time)) + dt);
}

template <typename T, int DPL>
__global__ void __launch_bounds__(256) k_solve(const int* __restrict__ gblob, int blob_words, MjtIn in, T* __restrict__ out, int B, int mode) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int E = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int env0 = blockIdx.x * E;
  const int nv = m.h(H_nv), nefc = m.h(H_nefc);
  const SolveLayout Lo = solve_layout(nv, nefc);
  T* region = reinterpret_cast<T*>(smem_blob + ((blob_words + 3) & ~3));

  // cooperative, env-contiguous loads of J, M, D, aref, frictionloss for the CTA's E envs
  auto load_tile = [&](int src_off, int rows, int cols, int dst_off, int ldd) {
    const int total = rows * cols * E;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int e = idx % E, k = idx / E;
      const int r = k / cols, cc = k - r * cols;
      const int env = env0 + e;
      if (env < B) region[(size_t)e * Lo.words + dst_off + r * ldd + cc] = out[(size_t)(src_off + k) * B + env];
    }
  };
  load_tile(m.D(D_qM), nv, nv, Lo.oM, Lo.ldm);
  if (nefc) {
    load_tile(m.D(D_efc_J), nefc, nv, Lo.oJ, Lo.ldj);
    load_tile(m.D(D_efc_D), 1, nefc, Lo.oD, nefc);
    load_tile(m.D(D_efc_aref), 1, nefc, Lo.oAref, nefc);
    load_tile(m.D(D_efc_frictionloss), 1, nefc, Lo.oFloss, nefc);
  }
  __syncthreads();
  const int env = env0 + warp;
  if (env >= B) return;
  EnvBuf<T> d{out + env, (size_t)B};
  solve_env<T, DPL>(m, in, d, (size_t)B, env, region + (size_t)warp * Lo.words, mode);
}
