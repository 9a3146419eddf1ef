// msv_launch.cu -- instantiations and launch wrappers of the SSV/MSV kernel family (launch.hpp)
#include <algorithm>

#include "launch.hpp"

namespace pa {

template <int Q>
static cudaError_t launch_msv(MsvArgs &a, int nrows_max, int num_sms, cudaStream_t st, uint64_t *launched) {
  const size_t smem = (size_t)nrows_max * Q * 32 * 4;
  cudaError_t e = cudaFuncSetAttribute(msv_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int bps = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, msv_kernel<Q>, 256, smem);
  bps = std::max(1, bps);
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)a.nlist * ntiles;
  const int grid = (int)std::min<unsigned long long>(nitems, (unsigned long long)num_sms * bps);
  if (grid > 0) { msv_kernel<Q><<<grid, 256, smem, st>>>(a); ++*launched; }
  return cudaGetLastError();
}

template <int Q, bool FT = false>
static cudaError_t launch_msv_tmem(MsvArgs &a, int nrows, int num_sms, cudaStream_t st, uint64_t *launched) {
  using Cfg = MsvTmemCfg<Q, FT>;
  constexpr int G = Cfg::G;
  const int need = msv_tmem_cols_needed(Q, nrows, FT);
  int ncols = 32;
  while (ncols < need) ncols *= 2;
  // resident CTAs per SM: fixed by the kernel's __launch_bounds__ (register cap) and by TMEM -- every
  // resident CTA must own its columns, or it would spin in tcgen05.alloc until a neighbour exits
  const int bps = std::max(1, std::min(Cfg::MINB, 512 / ncols));
  const size_t smem = (size_t)(nrows + 1) * Cfg::QS * 32 * 4;  // hybrid (Q > 16) only
  if (smem) {
    cudaError_t e = cudaFuncSetAttribute(msv_tmem_kernel<Q, G, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)a.nlist * ntiles;
  const int grid = (int)std::min<unsigned long long>(nitems, (unsigned long long)num_sms * bps);
  if (grid > 0) { msv_tmem_kernel<Q, G, FT><<<grid, Cfg::THREADS, smem, st>>>(a, ncols); ++*launched; }
  return cudaGetLastError();
}

cudaError_t launch_msv_q(int Q, bool tmem, bool fullt, MsvArgs &a, int nrows, int num_sms, cudaStream_t st, uint64_t *launched) {
  if (tmem && fullt) {
    switch (Q) {
#define C_(q) case q: return launch_msv_tmem<q, true>(a, nrows, num_sms, st, launched);
      C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26)
#undef C_
    }
    return cudaErrorInvalidValue;
  }
  if (tmem) {
    switch (Q) {
#define C_(q) case q: return launch_msv_tmem<q>(a, nrows, num_sms, st, launched);
      C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14) C_(15) C_(16)
      C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27) C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
    }
    return cudaErrorInvalidValue;
  }
  switch (Q) {
#define C_(q) case q: return launch_msv<q>(a, nrows, num_sms, st, launched);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14) C_(15) C_(16)
    C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27) C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_msv_pack(int QP, MsvPackArgs &pk, int ncols, int grid, cudaStream_t st, uint64_t *launched) {
  if (grid <= 0) return cudaSuccess;
  if (QP == 16) msv_pack_kernel<16, 2><<<grid, 1024, 0, st>>>(pk, ncols);
  else if (QP == PA_TMEM_FULLQ) msv_pack_kernel<PA_TMEM_FULLQ, 1><<<grid, 512, 0, st>>>(pk, ncols);
  else return cudaErrorInvalidValue;
  ++*launched;
  return cudaGetLastError();
}

cudaError_t launch_msv_wide(int QW, bool fullt, MsvWideArgs &wa, int ncols, int grid, size_t smem, cudaStream_t st, uint64_t *launched) {
  if (grid <= 0) return cudaSuccess;
  if (fullt) {
    switch (QW) {
#define C_(q)                                                \
  case q:                                                    \
    msv_wide_kernel<q, 1, true><<<grid, 512, 0, st>>>(wa, ncols); \
    ++*launched;                                             \
    return cudaGetLastError();
      C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26)
#undef C_
    }
    return cudaErrorInvalidValue;
  }
  switch (QW) {
#define C_(q)                                                                                                            \
  case q: {                                                                                                              \
    cudaError_t e = cudaFuncSetAttribute(msv_wide_kernel<q, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                      \
    msv_wide_kernel<q, 2><<<grid, 512, smem, st>>>(wa, ncols);                                                           \
    ++*launched;                                                                                                         \
    return cudaGetLastError();                                                                                           \
  }
    C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27) C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_msv_exact_wide(MsvExactWideArgs &ex, int num_sms, cudaStream_t st, uint64_t *launched) {
  int wpb = 4;
  const size_t per_warp = (size_t)ex.maxQ * 32 * 4;
  while (wpb > 1 && per_warp * wpb > 160 * 1024) wpb /= 2;
  const size_t smem = per_warp * wpb;
  cudaError_t e = cudaFuncSetAttribute(msv_exact_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int bps = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, msv_exact_wide_kernel, wpb * 32, smem);
  msv_exact_wide_kernel<<<num_sms * std::max(1, bps), wpb * 32, smem, st>>>(ex);
  ++*launched;
  return cudaGetLastError();
}

}  // namespace pa
