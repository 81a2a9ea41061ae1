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

__host__ __device__ inline SolveLayout solve_layout(int nv, int nefc, int jc_words, int nf, int nsrc, int tsize, bool compact, int G) {
  SolveLayout s;
  s.nP = nv * (nv + 1) / 2;
  s.nchunk = (nefc + G - 1) / G;   // one mask word per G rows (G = lanes per env)
  int o = 0;
  s.oJ = o; o += jc_words;
  s.oM = o; o += s.nP;
  s.oL = o; o += s.nP;
  s.oDinv = o; o += nv;
  s.oXs = o; o += nv;
  s.oYs = o; o += nv;
  s.oD = o; o += nefc;
  s.oAref = o; o += nefc;
  s.oFloss = o; o += nf > 0 ? nefc : 0;
  s.oJaref = o; o += nefc;
  s.oJv = o; o += nefc;
  s.oForce = o; o += nefc;
  s.oMask = o; o += s.nchunk;  // active-row bitmask over the *live* rows (uint32 per 32 rows, stored in T-sized slots)
  // live-row list (nefc ints) + compact position of every source (nsrc ints), packed as ints
  s.oLive = o; o += compact ? ((nefc + nsrc) * 4 + tsize - 1) / tsize : 0;
  s.words = (o + 3) & ~3;
  return s;
}


struct MjtIn;
// Launches k_solve<T, DPL, COMPACT, G> for the requested instantiation (lanes = G in {16, 32}, dpl in {1, 2}); returns a
// CUDA error code.  smem_attr is the largest dynamic shared memory this model can request (set once per function).
cudaError_t mjt_launch_solve_f32(int lanes, int dpl, bool compact, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st,
                                 const int* blob, const int* sblob, int solve_words, const MjtIn& in, float* out, int B, int mode);
cudaError_t mjt_launch_solve_f64(int lanes, int dpl, bool compact, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st,
                                 const int* blob, const int* sblob, int solve_words, const MjtIn& in, double* out, int B, int mode);
