// Solver launch shape constants, the per-env shared-memory layout of k_solve, and the launch entry points of the two
// solver translation units (mjt_solve_f32.cu / mjt_solve_f64.cu are compiled in parallel with mjt_api.cu).
#pragma once
#include <cuda_runtime.h>
// f64 launch shape of k_solve: threads per CTA (= 32 x envs per CTA) and minimum CTAs per SM (caps registers/thread)
#ifndef MJT_SOLVE_F64_THREADS
#define MJT_SOLVE_F64_THREADS 512
#endif
#ifndef MJT_CHOL_ROWS
#define MJT_CHOL_ROWS 1   // n <= 32: keep each lane's row of L in registers during the factorisation
#endif
#ifndef MJT_COMPACT_MIN_NEFC
#define MJT_COMPACT_MIN_NEFC 64   // models with more constraint rows than this use the live-row (COMPACT) solver instantiation
#endif
#ifndef MJT_SOLVE_F64_MINB
#define MJT_SOLVE_F64_MINB 1
#endif

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

struct SolveLayout {
  int nP, oJ, oM, oL, oDinv, oXs, oYs, oD, oAref, oFloss, oJaref, oJv, oForce, oMask, oLive, nchunk, words;
};

// rcap / jcap > 0 select the live tile: row vectors hold rcap row slots and the J tile jcap words (only sources that are
// live this step are loaded); 0 = every row / the whole compact J.
__host__ __device__ inline SolveLayout solve_layout(int nv, int nefc, int jc_words, int nf, int nsrc, int tsize, bool compact, int G,
                                                    int rcap = 0, int jcap = 0) {
  SolveLayout s;
  const int nrow = rcap > 0 ? rcap : nefc;
  s.nP = nv * (nv + 1) / 2;
  s.nchunk = (nrow + G - 1) / G;   // one mask word per G rows (G = lanes per env)
  int o = 0;
  s.oJ = o; o += rcap > 0 ? jcap : jc_words;
  s.oM = o; o += s.nP;
  s.oL = o; o += s.nP;
  s.oDinv = o; o += nv;
  s.oXs = o; o += nv;
  s.oYs = o; o += nv;
  s.oD = o; o += nrow;
  s.oAref = o; o += nrow;
  s.oFloss = o; o += nf > 0 ? nrow : 0;
  s.oJaref = o; o += nrow;
  s.oJv = o; o += nrow;
  s.oForce = o; o += nrow;
  s.oMask = o; o += s.nchunk;  // active-row bitmask over the *live* rows (one word per G rows, stored in T-sized slots)
  // ints: live-row list (nrow), compact position of every source (nsrc) and, for the live tile, its J offset (nsrc)
  s.oLive = o; o += compact ? ((nrow + nsrc * (rcap > 0 ? 2 : 1)) * 4 + tsize - 1) / tsize : 0;
  // odd region stride: the cooperative load / store phase has consecutive threads serve consecutive envs (one region apart),
  // which is conflict-free across the shared-memory banks only when the stride is odd in units of T
  s.words = o | 1;
  return s;
}

struct MjtIn;
// Launches k_solve<T, DPL, COMPACT, G> for the requested instantiation (lanes = G in {16, 32}, dpl in {1, 2}); returns a
// CUDA error code.  smem_attr is the largest dynamic shared memory this model can request (set once per function).
// compact: 0 = every row in place, 1 = live-row list over the full tile, 2 = live tile (rcap row slots, jcap J words);
// only_flagged: process only envs whose overflow flag (hidden Data field scr_ovf) was set by a compact == 2 launch.
cudaError_t mjt_launch_solve_f32(int lanes, int dpl, int compact, int rcap, int jcap, int only_flagged, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st,
                                 const int* blob, const int* sblob, int solve_words, const MjtIn& in, float* out, float* pub, int B, int mode);
cudaError_t mjt_launch_solve_f64(int lanes, int dpl, int compact, int rcap, int jcap, int only_flagged, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st,
                                 const int* blob, const int* sblob, int solve_words, const MjtIn& in, double* out, double* pub, int B, int mode);
