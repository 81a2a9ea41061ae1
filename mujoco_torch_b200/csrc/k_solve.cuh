// k_solve: one WARP per environment; E environments (warps) per CTA share one staged ModelBlob.
// Fuses the reference's factor_m + qacc_smooth solve (smooth.py:311-354, forward.py:222-228), the primal
// CG / Newton solver with exact line search (solver.py:244-553), the Euler-with-damping / ImplicitFast
// re-factorisations (forward.py:313-328, 404-416; derivative.py:22-68) and _advance (forward.py:255-310).
//
// Mapping: lane i owns dof i (and i+32 when nv > 32): every nv-vector (qacc, Ma, grad, search ...) lives in
// registers, one element per lane; dot products and costs are warp-shuffle reductions.  Per-env matrices live in
// shared memory: M and one factor workspace as packed lower triangles, and the constraint Jacobian COMPACTED to
// each row's static column support (the dofs a contact / limit can touch), which is 3x smaller than dense J for
// Humanoid and is what lets 16 f64 envs share one SM.  The Newton Hessian M + J^T D J is assembled only over active
// rows and only over their support (rank-k updates of the packed triangle).  Because a warp == an environment,
// every data-dependent branch (active rows, line-search bracketing, early exit) is warp-uniform: inactive rows and
// converged envs cost nothing -- the per-env early exit the reference's masked while_loop emulates.
#pragma once
#include "mjt_solve_layout.h"
#include "mjt_common.cuh"
#include "k_smooth.cuh"

// 4/8-byte asynchronous global->shared copy (LDGSTS): no staging register, so a thread keeps all its copies in flight
template <typename T>
__device__ __forceinline__ void cp_async_word(T* smem_dst, const T* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T, int DPL>
struct Vec {
  T v[DPL];
};

// lanes of one env: a full warp (G = 32) or a half warp (G = 16, two envs per warp for models with nv <= 16)
template <int G> __device__ __forceinline__ unsigned group_mask() {
  return G == 32 ? 0xffffffffu : (0xffffu << ((threadIdx.x & 16)));
}
template <typename T, int G>
__device__ __forceinline__ T group_sum(T v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o, G);
  return v;
}
template <typename T, int G>
__device__ __forceinline__ T vec_bcast(const Vec<T, 1>& x, int i, unsigned gm) { return __shfl_sync(gm, x.v[0], i & (G - 1), G); }
template <typename T, int G>
__device__ __forceinline__ T vec_bcast(const Vec<T, 2>& x, int i, unsigned gm) {
  const T a = __shfl_sync(gm, x.v[0], i & (G - 1), G), b = __shfl_sync(gm, x.v[1], i & (G - 1), G);
  return (i >= G) ? b : a;
}

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// A warp is an environment (G == 32), so every data-dependent decision of the solver is warp-uniform -- but the compiler
// cannot see that a value reduced by shuffles is the same on all lanes, treats the branch as divergent and wraps every
// shuffle behind it in a WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair plus BSSY / BSYNC reconvergence code.  Broadcasting the
// decision from lane 0 (one SHFL) makes the uniformity visible.  Half-warp envs (G == 16) do diverge: value passes through.
template <int G> __device__ __forceinline__ int uni(int v) { return G == 32 ? __shfl_sync(0xffffffffu, v, 0) : v; }
template <int G> __device__ __forceinline__ bool unib(bool v) { return G == 32 ? (__shfl_sync(0xffffffffu, v ? 1 : 0, 0) != 0) : v; }

// In-place Cholesky of a packed lower triangle for n <= G (one row per lane): lane i keeps row i of L in registers, so
// the inner product reads only row j from shared memory (a broadcast) instead of two strided rows.  A free function of
// scalars and pointers only: MJT_CHOL_CALL = 1 compiles it once per kernel (__noinline__; the unrolled body is ~1.5 k
// instructions and has three call sites) without forcing the solver context out of registers.
#ifndef MJT_CHOL_CALL
#define MJT_CHOL_CALL 0
#endif
#if MJT_CHOL_CALL
#define MJT_CHOL_INLINE __noinline__
#else
#define MJT_CHOL_INLINE __forceinline__
#endif
template <typename T, int G>
__device__ MJT_CHOL_INLINE void cholesky_rows_fn(T* A, T* di, int n, int lane, unsigned gm, const T* idle) {
  const bool big = n > 16;
  T r[G];
  const bool mine = lane < n;
  T* rowi = A + tri(mine ? lane : 0);
#pragma unroll
  for (int j = 0; j < G; ++j) {
    if (j < n) {
      const T* rowj = A + tri(j);
      // unconditional load (a conditional one compiles to a branch): lanes above the diagonal or beyond n read `idle`, a word of
      // the env's region that nobody writes during the factorisation (racecheck-clean); whatever they compute from it is
      // discarded by the selects below and never stored
      const T* src = (mine && lane >= j) ? rowi + j : idle;
      T acc = *src, acc2 = T(0);
#pragma unroll
      for (int k = 0; k + 1 < j; k += 2) {
        acc -= r[k] * rowj[k];
        acc2 -= r[k + 1] * rowj[k + 1];
      }
      if (j & 1) acc -= r[j - 1] * rowj[j - 1];
      acc += acc2;
      T sdiag = __shfl_sync(gm, acc, j, G);
      if (big) sdiag += T(1e-10); else sdiag = t_max(sdiag, T(1e-12));
      const T inv = t_rsqrt(sdiag);
      const T lij = (lane == j) ? sdiag * inv : acc * inv;
      r[j] = lane >= j ? lij : T(0);
      if (mine && lane >= j) rowi[j] = lij;   // column j of other rows is not read during this step: no hazard
      if (lane == 0) di[j] = inv;
      __syncwarp(gm);
    }
  }
}

// Out-of-line copy for the half-warp instantiations (nv <= 16: HalfCheetah, Hopper, Walker2d, Ant): their three call sites
// then share one body, which measured 4 % faster there (instruction fetch); the full-warp kernels are faster with it inlined.
template <typename T, int G>
__device__ __noinline__ void cholesky_rows_call(T* A, T* di, int n, int lane, unsigned gm, const T* idle) {
  cholesky_rows_fn<T, G>(A, di, n, lane, gm, idle);
}

template <typename T, int DPL, bool COMPACT, int G, bool TILE>
struct SolveCtx {
  const ModelView<T>& m;
  T* s;  // this env's shared-memory region
  SolveLayout L;
  int nv, nefc, lane;
  const int* row_src;
  const int* row_jadr;
  const int* src_nsup;
  const int* src_cadr;
  const int* src_row;
  const int* src_nrows;
  const int* sup_col;
  const int* tri_ab;
  const int* dof_sadr;
  const int* dof_sent;
  const int* row_order;
  const int* dof_ent;   // per (dof, source) entry, two words: J offset of the entry's column | ns << 20, first row | nr << 16
  unsigned gm_;   // shuffle / barrier mask of this env's G lanes (half-warp envs); use gmask()
  // A full warp per env: the mask is the compile-time constant 0xffffffff, which lets the compiler emit plain SHFL / WARPSYNC.ALL
  // instead of the mask-checking collective sequence it needs for a run-time mask.
  __device__ __forceinline__ unsigned gmask() const { return G == 32 ? 0xffffffffu : gm_; }
  // COMPACT: row loops walk the live list (host picks it for models with more than two 32-row chunks); otherwise every
  // row is visited in place with no indirection
  static constexpr bool compact = COMPACT;
  // TILE (live tile, implies COMPACT): the row vectors are indexed by live slot and the J tile holds only the live sources,
  // at dynamic offsets jpos[source]; otherwise vectors are indexed by row and J sits at its static compact offsets
  static constexpr bool tile = TILE;
  __device__ __forceinline__ int vi(int idx, int r) const { return TILE ? idx : r; }                 // index into D / aref / Jaref / jv / force
  __device__ __forceinline__ int* jpos() const { return spos() + m.h(H_nsrc); }
  __device__ __forceinline__ const T* jsrc(int sidx, int r0) const { return Jc() + (TILE ? jpos()[sidx] : row_jadr[r0]); }   // J block of a source
  __device__ __forceinline__ int nrowcap() const { return TILE ? m.h(H_live_rcap) : nefc; }
  __device__ __forceinline__ int row_at(int idx) const { return compact ? lrow()[idx] : idx; }

  __device__ __forceinline__ T* Jc() const { return s + L.oJ; }
  __device__ __forceinline__ T* Mp() const { return s + L.oM; }
  __device__ __forceinline__ T* Lp() const { return s + L.oL; }
  __device__ __forceinline__ T* dinv() const { return s + L.oDinv; }
  __device__ __forceinline__ T* xs() const { return s + L.oXs; }
  __device__ __forceinline__ T* ys() const { return s + L.oYs; }
  __device__ __forceinline__ unsigned* amask() const { return reinterpret_cast<unsigned*>(s + L.oMask); }
  __device__ __forceinline__ int* lrow() const { return reinterpret_cast<int*>(s + L.oLive); }          // live row ids, ascending
  __device__ __forceinline__ int* spos() const { return reinterpret_cast<int*>(s + L.oLive) + (TILE ? m.h(H_live_rcap) : nefc); }   // source -> first live slot, -1 = dead
  __device__ __forceinline__ bool slot_active(int k) const { return (amask()[k / G] >> (k % G)) & 1u; }

  // A source whose Jacobian block and reference accelerations are all exactly zero (inactive contact, limit or
  // equality: constraint.py zeroes those rows) can never carry force: J.x - aref stays 0.  Such rows are dropped from
  // every row loop; they keep Jaref = jv = force = 0.
  // Live tile: slots and J offsets of the sources the constraint kernel flagged live (hidden field scr_live), then their
  // D / aref / frictionloss rows and Jacobian blocks gathered from the dense Data fields.  Returns the number of live rows,
  // or -1 when they do not fit the tile (the env is then left to the full-tile launch).
  __device__ __forceinline__ int build_tile(const EnvBuf<T>& d, T* Dv, T* arefv, T* flossv) const {
    const int nsrc = m.h(H_nsrc), rcap = m.h(H_live_rcap), jcap = m.h(H_live_jcap);
    int* lr = lrow();
    int* sp = spos();
    int* jp = jpos();
    int base = 0, jbase = 0;
    for (int s0 = 0; s0 < nsrc; s0 += G) {
      const int sidx = s0 + lane;
      int nr = 0, r0 = 0, nw = 0;
      bool live = false;
      if (sidx < nsrc) {
        r0 = src_row[sidx];
        nr = src_nrows[sidx];
        nw = nr * src_nsup[sidx];
        live = d.get(m.D(D_scr_live) + sidx) != T(0);
      }
      const int cnt = live ? nr : 0, jw = live ? nw : 0;
      int incl = cnt, jincl = jw;
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {
        const int v = __shfl_up_sync(gmask(), incl, o, G), w = __shfl_up_sync(gmask(), jincl, o, G);
        if (lane >= o) { incl += v; jincl += w; }
      }
      const int start = base + incl - cnt, jstart = jbase + jincl - jw;
      if (sidx < nsrc) {
        sp[sidx] = live ? start : -1;
        jp[sidx] = jstart;
        if (start + cnt <= rcap)
          for (int e = 0; e < cnt; ++e) lr[start + e] = r0 + e;
      }
      base += __shfl_sync(gmask(), incl, G - 1, G);
      jbase += __shfl_sync(gmask(), jincl, G - 1, G);
    }
    __syncwarp(gmask());
    if (base > rcap || jbase > jcap) return -1;
    const int oD = m.D(D_efc_D), oA = m.D(D_efc_aref), oF = m.D(D_efc_frictionloss), oJ = m.D(D_efc_J);
    for (int idx = lane; idx < base; idx += G) {
      const int r = lr[idx];
      Dv[idx] = d.get(oD + r);
      arefv[idx] = d.get(oA + r);
      if (flossv) flossv[idx] = d.get(oF + r);
    }
    T* J = Jc();
    const bool jcompact = m.h(H_layout_state) != 0;   // state layout: efc_J already is the support-compacted form
    for (int idx = 0; idx < base; ++idx) {   // one live row at a time, lanes over its support columns
      const int r = lr[idx], sidx = row_src[r], ns = src_nsup[sidx];
      const int* col = sup_col + src_cadr[sidx];
      T* dst = J + jp[sidx] + (r - src_row[sidx]) * ns;
      for (int k = lane; k < ns; k += G) dst[k] = d.get(jcompact ? oJ + row_jadr[r] + k : oJ + r * nv + col[k]);
    }
    __syncwarp(gmask());
    return base;
  }
  // Returns the number of live rows (warp-uniform).  The context itself stays immutable so that it lives in registers.
  __device__ __forceinline__ int build_live(const T* aref) const {
    if (!compact) return nefc;   // few rows: every row stays in place (slot = row), nothing to build
    const int nsrc = m.h(H_nsrc);
    int* lr = lrow();
    int* sp = spos();
    int base = 0;
    for (int s0 = 0; s0 < nsrc; s0 += G) {
      const int sidx = s0 + lane;
      int nr = 0, r0 = 0;
      bool live = false;
      if (sidx < nsrc) {
        r0 = src_row[sidx];
        nr = src_nrows[sidx];
        const int nw = nr * src_nsup[sidx];
        const T* jp = jsrc(sidx, r0);
        int nz = 0;   // branch-free so that the loads pipeline
#pragma unroll 4
        for (int k = 0; k < nw; ++k) nz |= (jp[k] != T(0)) ? 1 : 0;
        for (int e = 0; e < nr; ++e) nz |= (aref[r0 + e] != T(0)) ? 1 : 0;
        live = nz != 0;
      }
      const int cnt = live ? nr : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {
        const int v = __shfl_up_sync(gmask(), incl, o, G);
        if (lane >= o) incl += v;
      }
      const int start = base + incl - cnt;
      if (sidx < nsrc) {
        sp[sidx] = live ? start : -1;
        for (int e = 0; e < cnt; ++e) lr[start + e] = r0 + e;
      }
      base += __shfl_sync(gmask(), incl, G - 1, G);
    }
    __syncwarp(gmask());
    return base;
  }

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
    return group_sum<T, G>(p, gmask());
  }
  __device__ __forceinline__ void to_smem(const Vec<T, DPL>& a, T* dst) const {
    __syncwarp(gmask());
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int i = lane + G * k;
      if (i < nv) dst[i] = a.v[k];
    }
    __syncwarp(gmask());
  }
  // y = A x for a symmetric matrix stored as a packed lower triangle
  __device__ __forceinline__ Vec<T, DPL> symv(const T* A, const Vec<T, DPL>& x) const {
    to_smem(x, xs());
    Vec<T, DPL> y = zero();
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      // every lane walks all nv columns (same trip count: no divergence, unrollable); entry (i, c) of the symmetric matrix is
      // word tri(i) + c of the packed triangle for c <= i, else tri(c) + i.  Lanes without a dof compute row 0 and drop it.
      const int i = lane + G * k, ii = i < nv ? i : 0, ti = tri(ii);
      const T* x = xs();
      T acc = 0, acc2 = 0;   // two independent chains (even / odd columns)
      int c = 0;
#pragma unroll 2
      for (; c + 1 < nv; c += 2) {
        acc += A[c <= ii ? ti + c : tri(c) + ii] * x[c];
        acc2 += A[c + 1 <= ii ? ti + c + 1 : tri(c + 1) + ii] * x[c + 1];
      }
      if (c < nv) acc += A[c <= ii ? ti + c : tri(c) + ii] * x[c];
      y.v[k] = i < nv ? acc + acc2 : T(0);
    }
    return y;
  }
  // out[r] = J[r,:] . x (- sub[r])   over each row's compact support
  __device__ __forceinline__ void jmul(const Vec<T, DPL>& x, T* out, const T* sub, int nlive) const {
    to_smem(x, xs());
    for (int idx = lane; idx < nlive; idx += G) {
      // rows in place: visited in the order of decreasing support size (model.py row_order), so that the 32 rows of one pass
      // have (nearly) the same trip count -- Humanoid: its 32 contact rows in one pass, the 21 one-column limit rows in the other
      const int r = compact ? row_at(idx) : row_order[idx];
      const int sidx = row_src[r];
      const int ns = src_nsup[sidx];
      const int* col = sup_col + src_cadr[sidx];
      const T* jr = TILE ? jsrc(sidx, 0) + (r - src_row[sidx]) * ns : Jc() + row_jadr[r];
      const T* x = xs();
      T acc = 0, acc2 = 0;
      int k = 0;
      for (; k + 1 < ns; k += 2) { acc += jr[k] * x[col[k]]; acc2 += jr[k + 1] * x[col[k + 1]]; }
      if (k < ns) acc += jr[k] * x[col[k]];
      acc += acc2;
      out[vi(idx, r)] = sub ? acc - sub[vi(idx, r)] : acc;
    }
    __syncwarp(gmask());
  }
  // y = J^T f as a per-dof gather: lane i walks the (source, support position) pairs of dof i; rows of a source are
  // stored back to back (ns entries each).  f must be exactly zero on rows that carry no force.
  __device__ __forceinline__ Vec<T, DPL> jtmul(const T* f) const {
    Vec<T, DPL> out;
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int i = lane + G * k;
      T acc = T(0);
      if (i < nv) {
        for (int q = dof_sadr[i], qe = dof_sadr[i + 1]; q < qe; ++q) {
          if (TILE) {
            const int ent = dof_sent[q], sidx = ent & 0xffff, pos = ent >> 16;
            const int cs = spos()[sidx];
            if (cs < 0) continue;
            const int r0 = src_row[sidx], nr = src_nrows[sidx], ns = src_nsup[sidx];
            const T* jp = jsrc(sidx, r0) + pos;
            const T* fp = f + cs;
            for (int e = 0; e < nr; ++e) acc += jp[e * ns] * fp[e];
          } else {
            // static J layout: everything about the entry comes from two packed words (no dependent table walk)
            if (compact && spos()[dof_sent[q] & 0xffff] < 0) continue;
            const int w0 = dof_ent[2 * q], w1 = dof_ent[2 * q + 1];
            const int ns = w0 >> 20, nr = w1 >> 16;
            const T* jp = Jc() + (w0 & 0xfffff);
            const T* fp = f + (w1 & 0xffff);
            for (int e = 0; e < nr; ++e) acc += jp[e * ns] * fp[e];
          }
        }
      }
      out.v[k] = acc;
    }
    return out;
  }
  __device__ __forceinline__ void cholesky_rows(T* A) const {
    if constexpr (G == 16 && !MJT_CHOL_CALL) cholesky_rows_call<T, G>(A, dinv(), nv, lane, gmask(), xs());
    else cholesky_rows_fn<T, G>(A, dinv(), nv, lane, gmask(), xs());
  }
  __device__ __forceinline__ void cholesky(T* A) const {
    if (DPL == 1 && MJT_CHOL_ROWS) { cholesky_rows(A); return; }
    const int n = nv;
    const bool big = n > 16;
    T* di = dinv();
    for (int j = 0; j < n; ++j) {
      const T* rowj = A + tri(j);
      T t[DPL];
#pragma unroll
      for (int p = 0; p < DPL; ++p) {
        const int i = j + lane + G * p;
        t[p] = T(0);
        if (i < n) {
          const T* rowi = A + tri(i);
          T acc = rowi[j];
#pragma unroll 4
          for (int k = 0; k < j; ++k) acc -= rowi[k] * rowj[k];
          t[p] = acc;
        }
      }
      T sdiag = __shfl_sync(gmask(), t[0], 0, G);
      if (big) sdiag += T(1e-10); else sdiag = t_max(sdiag, T(1e-12));
      const T inv = t_rsqrt(sdiag);
      const T ljj = sdiag * inv;
      __syncwarp(gmask());
#pragma unroll
      for (int p = 0; p < DPL; ++p) {
        const int i = j + lane + G * p;
        if (i < n) A[tri(i) + j] = (i == j) ? ljj : t[p] * inv;
      }
      if (lane == 0) di[j] = inv;
      __syncwarp(gmask());
    }
  }
  // x <- (L L^T)^-1 x with x distributed one element per lane (reference small_cholesky_solve :132-168)
  __device__ __forceinline__ Vec<T, DPL> cho_solve(const T* Lm, Vec<T, DPL> x) const {
    const int n = nv;
    const T* di = dinv();
    if constexpr (DPL == 1) {
      // One dof per lane.  Lane j keeps its entry UNSCALED while the sweep passes over it: step i broadcasts the owner's
      // product x_i * dinv_i, and the owner applies the same product to itself after the sweep -- the same arithmetic as
      // scaling in place at step i, without the per-step "am I the owner" select.  The update is one predicated load + FMA.
      const int j = lane;
      const bool mine = j < n;
      const T dl = mine ? di[j] : T(0);
      T xv = x.v[0];
      const T* rowp = Lm + tri(mine ? j : 0);   // row j of L
#pragma unroll 4
      for (int i = 0; i < n; ++i) {             // forward: L y = b
        const T yi = __shfl_sync(gmask(), xv * dl, i, G);
        if (mine && j > i) xv -= rowp[i] * yi;
      }
      xv *= dl;
      const T* colp = Lm + (mine ? j : 0);
#pragma unroll 4
      for (int i = n - 1; i >= 0; --i) {        // backward: L^T z = y
        const T zi = __shfl_sync(gmask(), xv * dl, i, G);
        if (j < i) xv -= colp[tri(i)] * zi;
      }
      xv *= dl;
      x.v[0] = xv;
      return x;
    } else {
      // Two dofs per lane (j = lane and lane + G): the same unscaled-owner scheme, with the sweep split at i = G so that the
      // broadcasting slot is known at compile time.  Step i costs a multiply, a shuffle and one predicated load + FMA per slot.
      static_assert(DPL == 2, "cho_solve handles one or two dofs per lane");
      const bool m0 = lane < n, m1 = lane + G < n;
      const T d0 = m0 ? di[lane] : T(0), d1 = m1 ? di[lane + G] : T(0);
      T x0 = x.v[0], x1 = x.v[1];
      const T* row0 = Lm + tri(m0 ? lane : 0);
      const T* row1 = Lm + tri(m1 ? lane + G : 0);
      const int nlo = n < G ? n : G;
#pragma unroll 4
      for (int i = 0; i < nlo; ++i) {            // forward, pivots in slot 0
        const T yi = __shfl_sync(gmask(), x0 * d0, i, G);
        if (m0 && lane > i) x0 -= row0[i] * yi;
        if (m1) x1 -= row1[i] * yi;
      }
      for (int i = G; i < n; ++i) {               // forward, pivots in slot 1
        const T yi = __shfl_sync(gmask(), x1 * d1, i - G, G);
        if (m1 && lane + G > i) x1 -= row1[i] * yi;
      }
      x0 *= d0;
      x1 *= d1;
      for (int i = n - 1; i >= G; --i) {          // backward, pivots in slot 1
        const T zi = __shfl_sync(gmask(), x1 * d1, i - G, G);
        const T* rowi = Lm + tri(i);
        if (m0) x0 -= rowi[lane] * zi;
        if (lane + G < i) x1 -= rowi[lane + G] * zi;
      }
#pragma unroll 4
      for (int i = nlo - 1; i >= 0; --i) {        // backward, pivots in slot 0
        const T zi = __shfl_sync(gmask(), x0 * d0, i, G);
        if (lane < i) x0 -= Lm[tri(i) + lane] * zi;
      }
      x.v[0] = x0 * d0;
      x.v[1] = x1 * d1;
      return x;
    }
  }
  // bits k0 .. k0+nr-1 (nr <= 32) of the active-slot mask (one mask word per G slots)
  __device__ __forceinline__ unsigned active_bits(int k0, int nr) const {
    if (G == 32) {
      const unsigned* am = amask();
      const int w = k0 >> 5, sh = k0 & 31;
      const unsigned lo = am[w], hi = (sh && w + 1 < L.nchunk) ? am[w + 1] : 0u;
      const unsigned v = __funnelshift_r(lo, hi, sh);
      return nr >= 32 ? v : (v & ((1u << nr) - 1u));
    }
    unsigned v = 0;
    for (int e = 0; e < nr; ++e) v |= (slot_active(k0 + e) ? 1u : 0u) << e;
    return v;
  }
  // Lp <- Mp + sum over active rows r of D_r J_r^T J_r, source by source over the packed support pairs
  __device__ __forceinline__ void hessian(const T* D) const {
    T* H = Lp();
    const T* M0 = Mp();
    __syncwarp(gmask());   // the previous factor may still be read by lagging lanes
    for (int p = lane; p < L.nP; p += G) H[p] = M0[p];
    __syncwarp(gmask());
    // sources that touch a single dof (joint limits, dof friction): one lane per dof adds their diagonal terms
    for (int i = lane; i < (m.h(H_nsrc1) ? nv : 0); i += G) {
      T acc = T(0);
      for (int q = dof_sadr[i], qe = dof_sadr[i + 1]; q < qe; ++q) {
        const int sidx = dof_sent[q] & 0xffff;
        if (src_nsup[sidx] != 1) continue;
        const int r0 = src_row[sidx], nr = src_nrows[sidx];
        const int cs = compact ? spos()[sidx] : r0;
        if (cs < 0) continue;
        const T* jp = jsrc(sidx, r0);
        const T* Ds = D + (TILE ? cs : r0);
        for (int e = 0; e < nr; ++e)
          if (slot_active(cs + e)) acc += Ds[e] * jp[e] * jp[e];
      }
      H[tri(i) + i] += acc;
    }
    __syncwarp(gmask());
    const int nsrc = m.h(H_nsrc);
    // Multi-dof sources, in ascending order.  A full warp per env picks the candidates 32 at a time (one lane per source tests
    // "more than one dof, and live", one ballot) instead of walking every source through a dependent chain of table loads: in a
    // contact-rich scene most of the candidate contacts are dead in any given step.
    unsigned pick = 0;
    for (int sidx = 0; sidx < nsrc; ++sidx) {
      if (G == 32 && compact) {   // (few-row models keep the plain loop: measured 2 % faster on Humanoid)
        if ((sidx & 31) == 0) {
          const int mine_s = sidx + lane;
          const bool ok = mine_s < nsrc && src_nsup[mine_s] != 1 && (!compact || spos()[mine_s] >= 0);
          pick = __ballot_sync(0xffffffffu, ok);
        }
        const unsigned rest = pick >> (sidx & 31);
        if (!rest) { sidx |= 31; continue; }          // nothing left in this group of 32
        sidx += __ffs(rest) - 1;                        // next picked source (warp-uniform: pick is a ballot)
      } else {
        if (src_nsup[sidx] == 1) continue;
      }
      const int r0 = src_row[sidx], nr = src_nrows[sidx];
      const int cs = compact ? spos()[sidx] : r0;
      if (!(G == 32 && compact) && cs < 0) continue;
      unsigned act = 0;
      if (nr <= 32) act = active_bits(cs, nr);
      else for (int e = 0; e < nr; ++e) act |= slot_active(cs + e) ? 1u : 0u;
      act = (unsigned)uni<G>((int)act);
      if (!act) continue;
      const int ns = src_nsup[sidx];
      const int* col = sup_col + src_cadr[sidx];
      const int npair = (ns * (ns + 1)) >> 1;
      const T* j0 = jsrc(sidx, r0);   // rows of one source are stored back to back, ns entries each
      const T* Ds = D + (TILE ? cs : r0);
      for (int p = lane; p < npair; p += G) {
        const int ab = tri_ab[p], a = ab & 0xffff, b = ab >> 16;
        T acc = 0;
        if (nr <= 32) {
          unsigned mk = act;
          while (mk) {
            const int e = __ffs(mk) - 1;
            mk &= mk - 1;
            const T* jr = j0 + e * ns;
            acc += Ds[e] * jr[a] * jr[b];
          }
        } else {
          for (int e = 0; e < nr; ++e) {
            if (!slot_active(cs + e)) continue;
            const T* jr = j0 + e * ns;
            acc += Ds[e] * jr[a] * jr[b];
          }
        }
        H[tri(col[a]) + col[b]] += acc;  // support columns are sorted ascending: col[a] >= col[b]
      }
      __syncwarp(gmask());
    }
  }
};

template <typename T, int DPL, bool COMPACT, int G, bool TILE>
__device__ void solve_env(const ModelView<T>& m, const MjtIn& in, const EnvBuf<T>& d, const EnvBuf<T>& pub, size_t B, int env,
                          const SolveCtx<T, DPL, COMPACT, G, TILE>& c, const Vec<T, DPL>& qfrc_smooth, const Vec<T, DPL>& qacc_smooth, int mode) {
  typedef Vec<T, DPL> V;
  const int lane = c.lane;
  const unsigned gm = c.gmask();
  const int nv = c.nv, nefc = c.nefc, flags = m.h(H_disableflags);
  const SolveLayout& Lo = c.L;
  T* const sm = c.s;
  T* const Mm = c.Mp();
  T* const Lm = c.Lp();
  T* const D = sm + Lo.oD;
  T* const aref = sm + Lo.oAref;
  T* const floss = sm + Lo.oFloss;
  T* const Jaref = sm + Lo.oJaref;
  T* const jv = sm + Lo.oJv;
  T* const force = sm + Lo.oForce;

  auto loadv = [&](int off) {
    V r;
#pragma unroll
    for (int k = 0; k < DPL; ++k) { const int i = lane + G * k; r.v[k] = i < nv ? d.get(off + i) : T(0); }
    return r;
  };
  // results go to the caller's buffer through the P table (layout.py); an absent field (offset -1) is not stored
  auto storev = [&](int off, const V& x) {
    if (off < 0) return;
#pragma unroll
    for (int k = 0; k < DPL; ++k) { const int i = lane + G * k; if (i < nv) pub.set(off + i, x.v[k]); }
  };
  const bool fixed = (mode & MJT_MODE_FIXED) != 0;   // step(..., fixed_iterations=True): no termination tests
  auto axpy = [&](const V& a, T s, const V& b) {
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

  V qacc = qacc_smooth;
  V qfrc_con = c.zero();

  if (nefc == 0) {
    // forward.py:390-396: solve() never runs; qfrc_constraint and qacc_warmstart keep the caller's values
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int i = lane + G * k;
      if (i < nv) {
        qfrc_con.v[k] = in_get<T>(in.qfrc_constraint, B, env, i);
        pub.set(m.P(P_qacc_warmstart) + i, in_get<T>(in.qacc_warmstart, B, env, i));
      }
    }
  } else {
    const int ne_nf = m.h(H_ne) + m.h(H_nf);
    const bool use_floss = m.h(H_nf) > 0 && !(flags & DSBL_FRICTIONLOSS);
    const bool newton = m.h(H_solver) == SOL_NEWTON;
    const T scale = m.opt(OPT_meaninertia) * T(nv > 1 ? nv : 1);
    const T tolerance = m.opt(OPT_tolerance), ls_tolerance = m.opt(OPT_ls_tolerance);
    const int iterations = m.h(H_iterations), ls_iterations = m.h(H_ls_iterations);

    V Ma, grad, Mgrad = c.zero(), search = c.zero();
    T gauss = 0, cost = t_inf<T>(), prev_cost = 0;
    if (c.compact && !TILE)   // dead rows are never visited again: give them their final zeros now
      for (int r = lane; r < nefc; r += G) { Jaref[r] = T(0); jv[r] = T(0); force[r] = T(0); }
    // rows that can carry force this step; the rest are exactly-zero rows
    const int nlive = uni<G>(TILE ? c.build_tile(d, D, aref, m.h(H_nf) > 0 ? floss : nullptr) : c.build_live(aref));
    if (TILE) {
      if (lane == 0) d.set(m.D(D_scr_ovf), nlive < 0 ? T(1) : T(0));
      if (nlive < 0) return;   // does not fit the live tile: the full-tile launch that follows solves this env
    }

    // reference _update_constraint :320-357 (cost only, or cost + forces + J^T f)
    auto update_constraint = [&](bool with_force) {
      T cpart = 0;
      for (int ch = 0; (ch * G) < nlive; ++ch) {
        const int idx = (ch * G) + lane;
        bool act = false;
        T fo = 0;
        if (idx < nlive) {
          const int r = c.row_at(idx), q = c.vi(idx, r);
          const T ja = Jaref[q], dr = D[q];
          act = (ja < T(0)) || (r < ne_nf);
          T ff = 0;
          if (use_floss) {
            const T fl = floss[q];
            if (fl > T(0)) {
              const T rr = T(1) / (dr + (dr == T(0) ? T(MJT_MINVAL) : T(0)));
              const bool lin_neg = ja <= -rr * fl, lin_pos = ja >= rr * fl;
              if (lin_neg) { act = false; ff = fl; cpart += -T(0.5) * rr * fl * fl - fl * ja; }
              if (lin_pos) { act = false; ff = -fl; cpart += -T(0.5) * rr * fl * fl + fl * ja; }
            }
          }
          fo = (act ? dr * -ja : T(0)) + ff;
          if (act) cpart += T(0.5) * dr * ja * ja;
          if (with_force) force[q] = fo;
        }
        const unsigned am = (__ballot_sync(gm, act) >> (threadIdx.x & 31 & ~(G - 1))) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u));
        if (with_force && lane == 0) c.amask()[ch] = am;
      }
      __syncwarp(gm);
      if (with_force) qfrc_con = c.jtmul(force);
      T g = 0;
#pragma unroll
      for (int k = 0; k < DPL; ++k) g += (Ma.v[k] - qfrc_smooth.v[k]) * (qacc.v[k] - qacc_smooth.v[k]);
      gauss = T(0.5) * group_sum<T, G>(g, gm);
      prev_cost = cost;
      cost = group_sum<T, G>(cpart, gm) + gauss;
    };

    auto compute_grad = [&]() {
#pragma unroll
      for (int k = 0; k < DPL; ++k) grad.v[k] = Ma.v[k] - qfrc_smooth.v[k] - qfrc_con.v[k];
    };
    // reference _update_gradient :359-376 (the M^-1 / H^-1 application)
    // The Hessian M + sum over active rows of D_r J_r^T J_r is a function of the active set alone (D and J are fixed during the
    // solve), so its factor in Lm stays valid for as long as the set does not change: lane k remembers word k of the mask the factor
    // was built from.  Lm starts as chol(M) from the factor_m prologue, i.e. the factor of the empty set.  Once the active set has
    // settled -- typically after a few Newton iterations of a resting contact scene -- the assembly and the factorisation are skipped.
    unsigned fmask = 0;
    const bool track = c.L.nchunk <= G;   // one mask word per lane; larger tiles always refactorise
    auto compute_mgrad = [&]() {
      if (newton) {
        const unsigned cur = lane < c.L.nchunk && (lane * G) < nlive ? c.amask()[lane] : 0u;
        const bool changed = !track || group_sum<int, G>(cur != fmask ? 1 : 0, gm) != 0;
        if (unib<G>(changed)) {
          c.hessian(D);
          c.cholesky(Lm);
          fmask = cur;
        }
      }
      Mgrad = c.cho_solve(Lm, grad);  // CG: Lm still holds chol(M)
    };

    // reference _linesearch :378-497
    auto linesearch = [&]() {
      const T smag = t_sqrt(c.dotv(search, search)) * scale;
      const T gtol = tolerance * ls_tolerance * smag;
      const V mv = c.symv(Mm, search);
      c.jmul(search, jv, nullptr, nlive);
      T qg0 = gauss, qg1, qg2;
      {
        T a = 0, b = 0, e = 0;
#pragma unroll
        for (int k = 0; k < DPL; ++k) { a += search.v[k] * Ma.v[k]; b += search.v[k] * qfrc_smooth.v[k]; e += search.v[k] * mv.v[k]; }
        qg1 = group_sum<T, G>(a, gm) - group_sum<T, G>(b, gm);
        qg2 = T(0.5) * group_sum<T, G>(e, gm);
      }
      struct Pt { T alpha, cost, d0, d1; };
      // evaluates NA step sizes in one pass over the constraint rows
      auto points = [&](const T* al, const int NA, Pt* out) {
        T q[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a][0] = q[a][1] = q[a][2] = T(0);
        for (int idx = lane; idx < nlive; idx += G) {
          const int r = c.row_at(idx), vq = c.vi(idx, r);
          const T ja = Jaref[vq], jr = jv[vq], dr = D[vq];
          const T c0 = T(0.5) * ja * ja * dr, c1 = jr * ja * dr, c2 = T(0.5) * jr * jr * dr;
          T fl = 0, rf = 0;
          const bool hasf = use_floss && (fl = floss[vq]) > T(0);
          if (hasf) rf = fl / (dr + (dr == T(0) ? T(MJT_MINVAL) : T(0)));
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (a < NA) {
              const T x = ja + al[a] * jr;
              bool act = (x < T(0)) || (r < ne_nf);
              if (hasf) {
                if (x <= -rf) { act = false; q[a][0] += fl * (-T(0.5) * rf - ja); q[a][1] += -fl * jr; }
                if (x >= rf) { act = false; q[a][0] += fl * (-T(0.5) * rf + ja); q[a][1] += fl * jr; }
              }
              if (act) { q[a][0] += c0; q[a][1] += c1; q[a][2] += c2; }
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (a < NA) {
            const T t0 = qg0 + group_sum<T, G>(q[a][0], gm), t1 = qg1 + group_sum<T, G>(q[a][1], gm), t2 = qg2 + group_sum<T, G>(q[a][2], gm);
            out[a].alpha = al[a];
            out[a].cost = al[a] * al[a] * t2 + al[a] * t1 + t0;
            out[a].d0 = T(2) * al[a] * t2 + t1;
            out[a].d1 = T(2) * t2 + (t2 == T(0) ? T(MJT_MINVAL) : T(0));
          }
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
        if (!fixed) {   // solver.py:484-487: the fixed loop runs ls_iterations bodies whatever the bracket looks like
          done = done || !swap;
          done = done || (lo.d0 < T(0) && lo.d0 > -gtol);
          done = done || (hi.d0 > T(0) && hi.d0 < gtol);
        }
        if (unib<G>(done)) break;
        Pt cand[3];
        al[0] = lo.alpha - lo.d0 / lo.d1;
        al[1] = hi.alpha - hi.d0 / hi.d1;
        al[2] = T(0.5) * (lo.alpha + hi.alpha);
        points(al, 3, cand);
        const Pt lo_next = cand[0], hi_next = cand[1], mid = cand[2];
        const bool not_bracketed = (lo.d0 < T(0)) == (hi.d0 < T(0));
        auto sw = [&](T cur, T cd) {
          const bool in_bracket = ((cur < cd) && (cd < T(0))) || ((cur > cd) && (cd > T(0)));
          return in_bracket || (not_bracketed && (t_abs(cd) < t_abs(cur)));
        };
        const bool s1 = sw(lo.d0, lo_next.d0); if (s1) lo = lo_next;
        const bool s2 = sw(lo.d0, mid.d0); if (s2) lo = mid;
        const bool s3 = sw(lo.d0, hi_next.d0); if (s3) lo = hi_next;
        const bool s4 = sw(hi.d0, hi_next.d0); if (s4) hi = hi_next;
        const bool s5 = sw(hi.d0, mid.d0); if (s5) hi = mid;
        const bool s6 = sw(hi.d0, lo_next.d0); if (s6) hi = lo_next;
        swap = s1 || s2 || s3 || s4 || s5 || s6;
        ++it;
      }
      const bool improved = unib<G>((lo.cost < p0.cost) || (hi.cost < p0.cost));
      const T alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
      if (improved) {
        qacc = axpy(qacc, alpha, search);
        Ma = axpy(Ma, alpha, mv);
        for (int idx = lane; idx < nlive; idx += G) { const int q = c.vi(idx, c.row_at(idx)); Jaref[q] += jv[q] * alpha; }
        __syncwarp(gm);
      }
    };

    // ---- warm start (:526-531): cost at qacc_warmstart vs cost at qacc_smooth; the winner's context is kept
    // instead of being rebuilt (same values: the reference's third _create_context repeats the same arithmetic)
    bool take_warm = false;
    V warm = c.zero(), Ma_w = c.zero();
    T cost_w = t_inf<T>();
    if (!(flags & DSBL_WARMSTART)) {
#pragma unroll
      for (int k = 0; k < DPL; ++k) { const int i = lane + G * k; warm.v[k] = i < nv ? in_get<T>(in.qacc_warmstart, B, env, i) : T(0); }
      qacc = warm;
      c.jmul(qacc, Jaref, aref, nlive);
      Ma = c.symv(Mm, qacc);
      Ma_w = Ma;
      cost = t_inf<T>();
      update_constraint(false);
      cost_w = cost;
    }
    qacc = qacc_smooth;
    c.jmul(qacc, jv, aref, nlive);  // Jaref(smooth) parked in jv while the two costs are compared
    Ma = c.symv(Mm, qacc);
    {
      T* keep = Jaref;
      // evaluate the smooth cost on the parked values
      T cpart = 0;
      for (int idx = lane; idx < nlive; idx += G) {
        const int r = c.row_at(idx), q = c.vi(idx, r);
        const T ja = jv[q], dr = D[q];
        bool act = (ja < T(0)) || (r < ne_nf);
        if (use_floss) {
          const T fl = floss[q];
          if (fl > T(0)) {
            const T rr = T(1) / (dr + (dr == T(0) ? T(MJT_MINVAL) : T(0)));
            if (ja <= -rr * fl) { act = false; cpart += -T(0.5) * rr * fl * fl - fl * ja; }
            if (ja >= rr * fl) { act = false; cpart += -T(0.5) * rr * fl * fl + fl * ja; }
          }
        }
        if (act) cpart += T(0.5) * dr * ja * ja;
      }
      const T cost_s = group_sum<T, G>(cpart, gm);  // gauss term is exactly zero at qacc_smooth
      take_warm = unib<G>(cost_w < cost_s);
      if (take_warm) {
        qacc = warm;
        Ma = Ma_w;
      } else {
        if (TILE) { for (int idx = lane; idx < nlive; idx += G) keep[idx] = jv[idx]; }
        else { for (int r = lane; r < nefc; r += G) keep[r] = jv[r]; }
      }
      __syncwarp(gm);
    }
    cost = t_inf<T>();
    prev_cost = 0;
    update_constraint(true);
    compute_grad();

    // ---- main loop (:501-539).  The gradient solve is only needed if an iteration follows, so the termination test is
    // evaluated first, here and at the end of body() (identical outputs, one factorisation less per solve).
    int niter = 0;
    bool run = true;
    if (iterations != 1 && !fixed) {
      const T improvement = (prev_cost - cost) / scale;
      const T gradient = t_sqrt(c.dotv(grad, grad)) / scale;
      run = unib<G>(!(niter >= iterations || improvement < tolerance || gradient < tolerance));
    }
    if (run) {
      compute_mgrad();
      search = neg(Mgrad);
    }
    // Who reads what the LAST body's _update_constraint produces (forces, J^T f, cost, gradient)?  The termination test -- unless the
    // iteration budget ends the loop anyway --, the efc_force / qfrc_constraint outputs (absent from the state layout) and the
    // integrator's velocity-implicit re-solve.  When none applies (Humanoid: 1 iteration, Euler without eulerdamp, state layout) the
    // step only needs qacc, which the line search has just updated: that last evaluation is skipped.
    const int integ_ = m.h(H_integrator);
    const bool resolve = (mode & MJT_MODE_KIND) == MJT_MODE_STEP &&
                         ((integ_ == INT_EULER && !(flags & DSBL_EULERDAMP)) || integ_ == INT_IMPLICITFAST);
    const bool need_final = resolve || m.P(P_qfrc_constraint) >= 0 || m.D(D_efc_force) >= 0;
    while (run) {
      linesearch();
      if (!need_final && niter + 1 >= iterations) { ++niter; break; }
      const V prev_grad = grad, prev_Mgrad = Mgrad;
      update_constraint(true);
      compute_grad();
      ++niter;
      if (iterations == 1) break;
      if (fixed) {   // solver.py:536-537: exactly `iterations` bodies
        if (niter >= iterations) break;
      } else {
        const T improvement = (prev_cost - cost) / scale;
        const T gradient = t_sqrt(c.dotv(grad, grad)) / scale;
        if (unib<G>(niter >= iterations || improvement < tolerance || gradient < tolerance)) break;
      }
      compute_mgrad();
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
    }
    storev(m.P(P_qacc_warmstart), qacc);
    if (m.D(D_efc_force) < 0) {
      // state layout: efc_force is not an output
    } else if (TILE) {   // dead rows carry no force; live rows scatter from their slots
      for (int r = lane; r < nefc; r += G) d.set(m.D(D_efc_force) + r, T(0));
      __syncwarp(gm);
      for (int idx = lane; idx < nlive; idx += G) d.set(m.D(D_efc_force) + c.lrow()[idx], force[idx]);
    } else {
      for (int r = lane; r < nefc; r += G) d.set(m.D(D_efc_force) + r, force[r]);
    }
  }
  storev(m.P(P_qacc), qacc);
  storev(m.P(P_qfrc_constraint), qfrc_con);

  const int integrator = m.h(H_integrator);
  const bool split = pub.p != d.p;   // state layout: the workspace holds the pre-integration working copies
  if ((mode & MJT_MODE_KIND) != MJT_MODE_STEP || integrator == INT_RK4) {   // RK4 is sequenced on the host from forward passes
    if (split) {   // forward(): the (sanitised / normalised) state and act_dot still have to reach the caller's buffer
      for (int i = lane; i < m.h(H_nq); i += G) pub.set(m.P(P_qpos) + i, d.get(m.D(D_qpos) + i));
      for (int i = lane; i < nv; i += G) pub.set(m.P(P_qvel) + i, d.get(m.D(D_qvel) + i));
      for (int i = lane; i < m.h(H_na); i += G) {
        pub.set(m.P(P_act) + i, d.get(m.D(D_act) + i));
        pub.set(m.P(P_act_dot) + i, d.get(m.D(D_act_dot) + i));
      }
      if (lane == 0) pub.set(m.P(P_time), d.get(m.D(D_time)));
    }
    return;
  }

  // ---- integrator: velocity-implicit re-solve (forward.py:313-328 / 404-416)
  const T dt = m.opt(OPT_timestep);
  V qacc_int = qacc;
  bool refactor = false;
  __syncwarp(gm);   // every lane is done reading the solver's factor before Lm is rebuilt below (racecheck-clean)
  if (integrator == INT_EULER && !(flags & DSBL_EULERDAMP)) {
    for (int p = lane; p < Lo.nP; p += G) Lm[p] = Mm[p];
    __syncwarp(gm);
    for (int i = lane; i < nv; i += G) Lm[tri(i) + i] += dt * m.F(F_dof_damping)[i];
    refactor = true;
  } else if (integrator == INT_IMPLICITFAST) {
    // qderiv = moment^T diag(biasprm2*[affine] + gainprm2*[affine]*ctrl) moment - diag(damping)  (derivative.py:22-68)
    const int nu = m.h(H_nu);
    const bool actuation = !(flags & DSBL_ACTUATION), damper = !(flags & DSBL_DAMPER);
    if (actuation || damper || m.h(H_ntendon)) {
      for (int p = lane; p < Lo.nP; p += G) Lm[p] = Mm[p];
      __syncwarp(gm);
      if (damper)
        for (int i = lane; i < nv; i += G) Lm[tri(i) + i] += dt * m.F(F_dof_damping)[i];
      __syncwarp(gm);
      if (actuation) {
        // d(force_u)/d(velocity) of actuator u (derivative.py:40-55)
        auto act_vel = [&](int u) {
          T vel = 0;
          if (m.I(I_act_biastype)[u] == BIAS_AFFINE) vel += m.F(F_act_biasprm)[10 * u + 2];
          if (m.I(I_act_gaintype)[u] == GAIN_AFFINE) {
            const T cu = m.I(I_act_dyntype)[u] != DYN_NONE ? d.get(m.D(D_act) + m.I(I_act_actadr)[u]) : in_get<T>(in.ctrl, B, env, u);
            vel += m.F(F_act_gainprm)[10 * u + 2] * cu;
          }
          return vel;
        };
        // a hinge / slide joint transmission has the single moment entry gear[0] at its dof: a diagonal update
        auto simple_dof = [&](int u) {
          const int trn = m.I(I_act_trntype)[u];
          if (trn != TRN_JOINT && trn != TRN_JOINTINPARENT) return -1;
          const int j = m.I(I_act_trnid)[u], jt = m.I(I_jnt_type)[j];
          return (jt == JNT_HINGE || jt == JNT_SLIDE) ? m.I(I_jnt_dofadr)[j] : -1;
        };
        for (int u0 = 0; u0 < nu; u0 += G) {   // lane u evaluates actuator u; the updates are applied one by one
          const int u = u0 + lane;
          int dof = -1;
          T val = 0;
          if (u < nu) {
            dof = simple_dof(u);
            if (dof >= 0) { const T g = m.F(F_act_gear)[6 * u]; val = dt * g * act_vel(u) * g; }
          }
          const int cnt = (nu - u0) < G ? (nu - u0) : G;
          for (int k = 0; k < cnt; ++k) {
            const int dk = __shfl_sync(gm, dof, k, G);
            const T vk = __shfl_sync(gm, val, k, G);
            if (lane == 0 && dk >= 0 && vk != T(0)) Lm[tri(dk) + dk] -= vk;
          }
        }
        __syncwarp(gm);
        for (int u = 0; u < nu; ++u) {   // ball / free / tendon transmissions: dense rank-one update from the stored moment row
          if (simple_dof(u) >= 0) continue;
          const T vel = act_vel(u);
          if (vel == T(0)) continue;
          for (int p = lane; p < Lo.nP; p += G) {
            const int ab = c.tri_ab[p], i = ab & 0xffff, j = ab >> 16;
            Lm[p] -= dt * d.get(m.D(D_actuator_moment) + u * nv + i) * vel * d.get(m.D(D_actuator_moment) + u * nv + j);
          }
          __syncwarp(gm);
        }
      }
      for (int t = 0; t < m.h(H_ntendon); ++t) {
        const T td = m.F(F_ten_damping)[t];
        if (td == T(0)) continue;
        for (int p = lane; p < Lo.nP; p += G) {
          const int ab = c.tri_ab[p], i = ab & 0xffff, j = ab >> 16;
          Lm[p] += dt * d.get(m.D(D_ten_J) + t * nv + i) * td * d.get(m.D(D_ten_J) + t * nv + j);
        }
        __syncwarp(gm);
      }
      refactor = true;
    }
  }
  if (refactor) {
    __syncwarp(gm);
    c.cholesky(Lm);
    V rhs;
#pragma unroll
    for (int k = 0; k < DPL; ++k) rhs.v[k] = qfrc_smooth.v[k] + qfrc_con.v[k];
    qacc_int = c.cho_solve(Lm, rhs);
  }

  // ---- _advance (forward.py:255-310)
  const int nu = m.h(H_nu);
  for (int u = lane; u < nu; u += G) {
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
      pub.set(m.P(P_act) + a, nx);
    }
  }
  if (split) {   // activations without an owning actuator keep their value; act_dot is an output of the step
    const int* actadr = m.I(I_act_actadr);
    for (int a = lane; a < m.h(H_na); a += G) {
      bool owned = false;
      for (int u = 0; u < nu; ++u) owned |= actadr[u] >= 0 && a >= actadr[u] && a < actadr[u] + m.I(I_act_actnum)[u];
      if (!owned) pub.set(m.P(P_act) + a, d.get(m.D(D_act) + a));
      pub.set(m.P(P_act_dot) + a, d.get(m.D(D_act_dot) + a));
    }
  }
  V qvel = loadv(m.D(D_qvel));
  qvel = axpy(qvel, dt, qacc_int);
  storev(m.P(P_qvel), qvel);
  c.to_smem(qvel, c.xs());
  const T* nvel = c.xs();
  const int oQ = m.D(D_qpos), pQ = m.P(P_qpos);
  for (int j = lane; j < m.h(H_njnt); j += G) {
    const int jt = m.I(I_jnt_type)[j], qa = m.I(I_jnt_qposadr)[j], da = m.I(I_jnt_dofadr)[j];
    if (jt == JNT_FREE || jt == JNT_BALL) {
      int qq = qa, dd = da;
      if (jt == JNT_FREE) {
        for (int k = 0; k < 3; ++k) pub.set(pQ + qa + k, d.get(oQ + qa + k) + dt * nvel[da + k]);
        qq = qa + 3;
        dd = da + 3;
      }
      Q4<T> q(d.get(oQ + qq), d.get(oQ + qq + 1), d.get(oQ + qq + 2), d.get(oQ + qq + 3));
      q = qintegrate(q, V3<T>(nvel[dd], nvel[dd + 1], nvel[dd + 2]), dt);
      pub.set(pQ + qq, q.w); pub.set(pQ + qq + 1, q.x); pub.set(pQ + qq + 2, q.y); pub.set(pQ + qq + 3, q.z);
    } else {
      pub.set(pQ + qa, d.get(oQ + qa) + dt * nvel[da]);
    }
  }
  if (lane == 0) pub.set(m.P(P_time), d.get(m.D(D_time)) + dt);
}

template <typename T, int DPL, bool COMPACT, int G, bool TILE>
__global__ void __launch_bounds__((sizeof(T) == 4 ? 512 : MJT_SOLVE_F64_THREADS), (sizeof(T) == 4 ? (DPL == 1 ? 2 : 1) : MJT_SOLVE_F64_MINB)) k_solve(const int* __restrict__ gfull, const int* __restrict__ gblob, int blob_words, MjtIn in,
                                               T* __restrict__ out, T* __restrict__ pubbuf, int B, int mode, int only_flagged) {
  extern __shared__ __align__(16) int smem_blob[];
  __shared__ uint64_t bar;
  if (only_flagged) {
    // second pass after a live-tile launch: only envs whose tile overflowed (scr_ovf != 0) are solved here; a CTA without
    // any leaves before it stages anything
    const int e_cta = blockDim.x / G, e0 = blockIdx.x * e_cta;
    const int o_ovf = gblob[MJT_OFF_D + D_scr_ovf];
    bool mine = false;
    for (int e = threadIdx.x; e < e_cta; e += blockDim.x)
      if (e0 + e < B) mine = mine || (out[(size_t)o_ovf * (size_t)B + e0 + e] != T(0));
    if (!__syncthreads_or(mine ? 1 : 0)) return;
  }
  stage_blob(smem_blob, &bar, gblob, blob_words);
  const ModelView<T> m = make_view<T>(smem_blob);
  const int E = blockDim.x / G;                                   // envs per CTA, G lanes each
  // "warp" = env slot inside the CTA; for full-warp envs the broadcast tells the compiler the slot (hence `valid`, the region
  // pointer and every branch on per-env data below) is warp-uniform
  const int warp = uni<G>(threadIdx.x / G), lane = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const int env0 = blockIdx.x * E;
  const int nv = m.h(H_nv), nefc = m.h(H_nefc);
  const SolveLayout Lo = solve_layout(nv, nefc, m.h(H_jc_words), m.h(H_nf), m.h(H_nsrc), (int)sizeof(T), COMPACT, G,
                                      TILE ? m.h(H_live_rcap) : 0, TILE ? m.h(H_live_jcap) : 0);
  T* region = reinterpret_cast<T*>(smem_blob + ((blob_words + 3) & ~3));
  const size_t Bs = (size_t)B;

  // ---- cooperative, env-contiguous loads: thread t serves env t % E, so E consecutive envs share every word's cache line
  const int ce = threadIdx.x % E, ck0 = threadIdx.x / E, cks = blockDim.x / E;
  const bool cvalid = env0 + ce < B && (!only_flagged || out[(size_t)m.D(D_scr_ovf) * Bs + env0 + ce] != T(0));
  T* const crg = region + (size_t)ce * Lo.words;
  T* const cout = out + env0 + ce;
  const int* tri_ab = m.I(I_tri_ab);   // packed lower index p -> (a, b): dense word a*nv+b, mirrored word b*nv+a
  const bool lstate = m.h(H_layout_state) != 0;   // state layout: qM is a packed triangle, efc_J the support-compacted rows
  {
    // M: only the entries that can be non-zero (kinematic-tree pattern, model.py m_nz) are fetched; the rest of the tile is zeroed
    const int oM = m.D(D_qM), nmz = m.h(H_n_mnz);
    const int* mnz = m.I(I_m_nz);
    if (nmz < Lo.nP) {
      if (cvalid)
        for (int p = ck0; p < Lo.nP; p += cks) crg[Lo.oM + p] = T(0);
      __syncthreads();
    }
    if (cvalid)
      for (int q = ck0; q < nmz; q += cks) {
        const int p = mnz[q], ab = tri_ab[p];
        cp_async_word(crg + Lo.oM + p, cout + (size_t)(lstate ? oM + p : oM + (ab & 0xffff) * nv + (ab >> 16)) * Bs);
      }
    if (nefc && !TILE) {   // the live tile gathers its rows per env once it knows which sources are live
      const int oJ = m.D(D_efc_J), jcw = m.h(H_jc_words);
      const int* jsrc = m.I(I_jc_src);   // compact J entry -> dense word row*nv+col
      if (cvalid)
        for (int k = ck0; k < jcw; k += cks) cp_async_word(crg + Lo.oJ + k, cout + (size_t)(oJ + (lstate ? k : jsrc[k])) * Bs);
      const int oD = m.D(D_efc_D), oA = m.D(D_efc_aref), oF = m.D(D_efc_frictionloss);
      const bool hasf = m.h(H_nf) > 0;
      if (cvalid)
        for (int k = ck0; k < nefc; k += cks) {
          cp_async_word(crg + Lo.oD + k, cout + (size_t)(oD + k) * Bs);
          cp_async_word(crg + Lo.oAref + k, cout + (size_t)(oA + k) * Bs);
          if (hasf) cp_async_word(crg + Lo.oFloss + k, cout + (size_t)(oF + k) * Bs);
        }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int env = env0 + warp;
  const bool valid = env < B && (!only_flagged || out[(size_t)m.D(D_scr_ovf) * Bs + (env < B ? env : 0)] != T(0));
  EnvBuf<T> d{out + (valid ? env : 0), Bs};
  EnvBuf<T> pub{pubbuf + (valid ? env : 0), Bs};   // the caller's buffer (same as `out` in the full layout)
  SolveCtx<T, DPL, COMPACT, G, TILE> c{m, region + (size_t)warp * Lo.words, Lo, nv, nefc, lane,
                     m.I(I_row_src), m.I(I_row_jadr), m.I(I_src_nsup), m.I(I_src_cadr), m.I(I_src_row), m.I(I_src_nrows), m.I(I_sup_col), m.I(I_tri_ab), m.I(I_dof_sadr), m.I(I_dof_sent), m.I(I_row_order), m.I(I_dof_ent), group_mask<G>()};
  Vec<T, DPL> qfrc_smooth = c.zero(), qacc_smooth = c.zero();
  if (valid) {
    // ---- factor_m + qacc_smooth
    T* Lm = c.Lp();
    const T* Mm = c.Mp();
    for (int p = lane; p < Lo.nP; p += G) Lm[p] = Mm[p];
    __syncwarp(gm);
    c.cholesky(Lm);
#pragma unroll
    for (int k = 0; k < DPL; ++k) { const int i = lane + G * k; qfrc_smooth.v[k] = i < nv ? d.get(m.D(D_qfrc_smooth) + i) : T(0); }
    qacc_smooth = c.cho_solve(Lm, qfrc_smooth);
    if (m.D(D_qacc_smooth) >= 0) {
#pragma unroll
      for (int k = 0; k < DPL; ++k) { const int i = lane + G * k; if (i < nv) d.set(m.D(D_qacc_smooth) + i, qacc_smooth.v[k]); }
    }
  }
  if (m.D(D_qLD) >= 0) {  // cooperative store of qLD = chol(M) (dense lower, zeros above the diagonal); not an output of the state layout
    __syncthreads();
    const int oL = m.D(D_qLD);
    if (cvalid) {
      for (int p = ck0; p < Lo.nP; p += cks) {
        const int ab = tri_ab[p], a = ab & 0xffff, b = ab >> 16;
        cout[(size_t)(oL + a * nv + b) * Bs] = crg[Lo.oL + p];
        if (a != b) cout[(size_t)(oL + b * nv + a) * Bs] = T(0);
      }
    }
    __syncthreads();
  }
  if (!valid) return;
  solve_env<T, DPL, COMPACT, G, TILE>(m, in, d, pub, Bs, env, c, qfrc_smooth, qacc_smooth, mode);
}
