// k_solve instantiations for float: compiled as its own translation unit so that the build runs in parallel.
#include "k_solve.cuh"

template <int DPL, bool COMPACT, int G, bool TILE>
static cudaError_t launch(int only_flagged, int grid, int threads, size_t smem, size_t smem_attr, cudaStream_t st, const int* blob,
                          const int* sblob, int solve_words, const MjtIn& in, float* out, float* pub, int B, int mode) {
  if (smem_attr > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute((const void*)k_solve<float, DPL, COMPACT, G, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attr);
    if (e != cudaSuccess) return e;
  }
  k_solve<float, DPL, COMPACT, G, TILE><<<grid, threads, smem, st>>>(blob, sblob, solve_words, in, out, pub, B, mode, only_flagged);
  return cudaGetLastError();
}

cudaError_t mjt_launch_solve_f32(int lanes, int dpl, int compact, int rcap, int jcap, int only_flagged, int grid, int threads, size_t smem,
                                 size_t smem_attr, cudaStream_t st, const int* blob, const int* sblob, int solve_words, const MjtIn& in,
                                 float* out, float* pub, int B, int mode) {
  (void)rcap; (void)jcap;   // the kernels read the caps from the staged blob header
#define MJT_CASE(D, GV)                                                                                                         \
  do {                                                                                                                          \
    if (compact == 2) return launch<D, true, GV, true>(only_flagged, grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, pub, B, mode);   \
    if (compact == 1) return launch<D, true, GV, false>(only_flagged, grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, pub, B, mode);  \
    return launch<D, false, GV, false>(only_flagged, grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, pub, B, mode);                   \
  } while (0)
  if (lanes == 16 && dpl == 1) MJT_CASE(1, 16);
  if (lanes == 16 && dpl == 2) MJT_CASE(2, 16);
  if (lanes == 32 && dpl == 1) MJT_CASE(1, 32);
  if (lanes == 32 && dpl == 2) MJT_CASE(2, 32);
#undef MJT_CASE
  return cudaErrorInvalidValue;
}
