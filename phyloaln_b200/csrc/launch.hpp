// launch.hpp -- launch wrappers of the templated kernel families, one translation unit per family
// (msv_launch.cu, vit_launch.cu, fwd_launch.cu) so that the ~250 instantiations build in parallel.
// Every function launches at most one kernel on `st` and returns cudaGetLastError(); *launched is
// incremented when a kernel was launched (the caller's gpu_launches counter).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "msv_tmem.cuh"
#include "vit16.cuh"
#include "wide.cuh"

namespace pa {

// msv_kernel<Q> (shared-memory DPX kernel) or msv_tmem_kernel<Q,2> (table in tensor memory; a.tab must point at the fp16 pool)
// fullt: rows entirely in TMEM (Q 17..26 for small alphabets, msv_tmem_full_ok)
cudaError_t launch_msv_q(int Q, bool tmem, bool fullt, MsvArgs &a, int nrows, int num_sms, cudaStream_t st, uint64_t *launched);
cudaError_t launch_msv_pack(int QP, MsvPackArgs &pk, int ncols, int grid, cudaStream_t st, uint64_t *launched);
cudaError_t launch_msv_wide(int QW, bool fullt, MsvWideArgs &wa, int ncols, int grid, size_t smem, cudaStream_t st, uint64_t *launched);
cudaError_t launch_msv_exact_wide(MsvExactWideArgs &ex, int num_sms, cudaStream_t st, uint64_t *launched);

// NW == 1: vit16_kernel<BH> / vit_kernel<Bv>; NW > 1: the multi-warp kernels of wide.cuh
cudaError_t launch_vit16_class(int BH, int NW, Vit16Args &v, int num_sms, cudaStream_t st, uint64_t *launched);
cudaError_t launch_vit_class(int Bv, int NW, VitArgs &v, int num_sms, cudaStream_t st, uint64_t *launched);

// B <= 64: fwd_kernel_t<B>.  B > 64: fwd_wide_kernel, one launch per size class (fwd_wide_class: what sets the shared-memory
// rows and so the CTAs per SM); `scratch` holds fwd_wide_grid_cap(num_sms) rows of fwd_wide_per_cta(maxB) floats
cudaError_t launch_fwd_b(int B, FwdTArgs &a, int num_sms, float *scratch, int maxB, cudaStream_t st, uint64_t *launched);
inline int fwd_wide_class(int B) { return B <= 64 ? B : B <= 96 ? 96 : B <= 128 ? 128 : B <= 192 ? 192 : 256; }
inline size_t fwd_wide_per_cta(int maxB) { return (size_t)maxB * 32; }
inline int fwd_wide_grid_cap(int num_sms) { return num_sms * 8; }

}  // namespace pa
