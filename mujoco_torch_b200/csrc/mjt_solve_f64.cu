// k_solve instantiations for double: compiled as its own translation unit so that the build runs in parallel.
#include "k_solve.cuh"

template <int DPL, bool COMPACT, int G>
static cudaError_t launch(int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st, const int* blob, const int* sblob,
                          int solve_words, const MjtIn& in, double* out, int B, int mode) {
  if (smem_attr > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute((const void*)k_solve<double, DPL, COMPACT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attr);
    if (e != cudaSuccess) return e;
  }
  k_solve<double, DPL, COMPACT, G><<<grid, threads, smem, st>>>(blob, sblob, solve_words, in, out, B, mode);
  return cudaGetLastError();
}

cudaError_t mjt_launch_solve_f64(int lanes, int dpl, bool compact, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st,
                                 const int* blob, const int* sblob, int solve_words, const MjtIn& in, double* out, int B, int mode) {
#define MJT_CASE(D, C, GV) return launch<D, C, GV>(grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, B, mode)
  if (lanes == 16 && dpl == 1) { if (compact) MJT_CASE(1, true, 16); else MJT_CASE(1, false, 16); }
  if (lanes == 16 && dpl == 2) { if (compact) MJT_CASE(2, true, 16); else MJT_CASE(2, false, 16); }
  if (lanes == 32 && dpl == 1) { if (compact) MJT_CASE(1, true, 32); else MJT_CASE(1, false, 32); }
  if (lanes == 32 && dpl == 2) { if (compact) MJT_CASE(2, true, 32); else MJT_CASE(2, false, 32); }
#undef MJT_CASE
  return cudaErrorInvalidValue;
}
