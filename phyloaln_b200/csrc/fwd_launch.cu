// fwd_launch.cu -- instantiations and launch wrappers of the Forward filter kernels (launch.hpp)
#include <algorithm>

#include "launch.hpp"

namespace pa {

template <int B>
static cudaError_t launch_fwd_t(FwdTArgs &a, int num_sms, cudaStream_t st, uint64_t *launched) {
  const unsigned long long n = a.end - a.begin;
  if (n == 0) return cudaSuccess;
  // B <= 32: rows in registers, 4 warps per block. B > 32: 4 shared-memory rows per warp, as many warps as fit in ~96 KB
  constexpr size_t per_warp = B <= 32 ? 0 : (size_t)4 * B * 32 * 4;
  int wpb = 4;
  if constexpr (per_warp != 0) wpb = (int)std::max<size_t>(1, std::min<size_t>(4, (96 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  if (smem) {
    cudaError_t e = cudaFuncSetAttribute(fwd_kernel_t<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = (int)std::min<unsigned long long>((n + wpb - 1) / wpb, (unsigned long long)num_sms * 32);
  fwd_kernel_t<B><<<grid, wpb * 32, smem, st>>>(a);
  ++*launched;
  return cudaGetLastError();
}

cudaError_t launch_fwd_b(int B, FwdTArgs &a, int num_sms, float *scratch, int maxB, cudaStream_t st, uint64_t *launched) {
  if (B > 64) {  // wide profiles: run-time block size, rows of 3 x class x 32 floats in shared memory, one warp per CTA
    const unsigned long long n = a.end - a.begin;
    if (n == 0) return cudaSuccess;
    if (!scratch) return cudaErrorInvalidValue;
    const size_t smem = (size_t)3 * fwd_wide_class(B) * 32 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fwd_wide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd_wide_kernel, 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const unsigned long long cap = (unsigned long long)std::min(fwd_wide_grid_cap(num_sms), num_sms * per_sm);
    FwdWideArgs w;
    w.ft = a;
    w.per_cta = fwd_wide_per_cta(maxB);
    w.scratch = scratch;
    fwd_wide_kernel<<<(unsigned)std::min(n, cap), 32, smem, st>>>(w);
    ++*launched;
    return cudaGetLastError();
  }
  switch (B) {
#define C_(b) case b: return launch_fwd_t<b>(a, num_sms, st, launched);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14) C_(15) C_(16)
    C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27) C_(28) C_(29) C_(30) C_(31) C_(32)
    C_(33) C_(34) C_(35) C_(36) C_(37) C_(38) C_(39) C_(40) C_(41) C_(42) C_(43) C_(44) C_(45) C_(46) C_(47) C_(48)
    C_(49) C_(50) C_(51) C_(52) C_(53) C_(54) C_(55) C_(56) C_(57) C_(58) C_(59) C_(60) C_(61) C_(62) C_(63) C_(64)
#undef C_
  }
  return cudaErrorInvalidValue;
}

}  // namespace pa
