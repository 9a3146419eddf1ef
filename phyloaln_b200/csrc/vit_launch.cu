// vit_launch.cu -- instantiations and launch wrappers of the Viterbi filter kernels (launch.hpp)
#include <algorithm>

#include "launch.hpp"

namespace pa {

template <int BT>
static cudaError_t launch_vit(VitArgs &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  const unsigned long long n = v.end - v.begin;
  if (n == 0) return cudaSuccess;
  const int grid = (int)std::min<unsigned long long>((n + 3) / 4, (unsigned long long)num_sms * 16);
  vit_kernel<BT><<<grid, 128, 0, st>>>(v);
  ++*launched;
  return cudaGetLastError();
}
template <int BH>
static cudaError_t launch_vit16(Vit16Args &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  const unsigned long long n = v.end - v.begin;
  if (n == 0) return cudaSuccess;
  const int grid = (int)std::min<unsigned long long>((n + 3) / 4, (unsigned long long)num_sms * 24);
  vit16_kernel<BH><<<grid, 128, 0, st>>>(v);
  ++*launched;
  return cudaGetLastError();
}
// wide profiles (wide.cuh): one pair per CTA of NW warps
template <int BH>
static cudaError_t launch_vit16_wide(int NW, Vit16Args &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  const unsigned long long n = v.end - v.begin;
  if (n == 0) return cudaSuccess;
  const int grid = (int)std::min<unsigned long long>(n, (unsigned long long)num_sms * (NW <= 2 ? 8 : 4));
  vit16_wide_kernel<BH><<<grid, NW * 32, 0, st>>>(v);
  ++*launched;
  return cudaGetLastError();
}
template <int BT>
static cudaError_t launch_vit_wide(int NW, VitArgs &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  const unsigned long long n = v.end - v.begin;
  if (n == 0) return cudaSuccess;
  const int grid = (int)std::min<unsigned long long>(n, (unsigned long long)num_sms * (NW <= 4 ? 2 : 1));
  vit_wide_kernel<BT><<<grid, NW * 32, 0, st>>>(v);
  ++*launched;
  return cudaGetLastError();
}

cudaError_t launch_vit16_class(int BH, int NW, Vit16Args &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  if (NW > 1) {
    switch (BH) {
#define C_(b) case b: return launch_vit16_wide<b>(NW, v, num_sms, st, launched);
      C_(16) C_(20) C_(24) C_(28) C_(32)
#undef C_
    }
    return cudaErrorInvalidValue;
  }
  switch (BH) {
#define C_(b) case b: return launch_vit16<b>(v, num_sms, st, launched);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(10) C_(12) C_(14) C_(16) C_(20) C_(24) C_(28) C_(32)
#undef C_
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_vit_class(int Bv, int NW, VitArgs &v, int num_sms, cudaStream_t st, uint64_t *launched) {
  if (NW > 1) {
    switch (Bv) {
#define C_(b) case b: return launch_vit_wide<b>(NW, v, num_sms, st, launched);
      C_(16) C_(20) C_(24) C_(28) C_(32)
#undef C_
    }
    return cudaErrorInvalidValue;
  }
  switch (Bv) {
#define C_(b) case b: return launch_vit<b>(v, num_sms, st, launched);
    C_(1) C_(2) C_(3) C_(4) C_(6) C_(8) C_(10) C_(12) C_(14) C_(16) C_(20) C_(24) C_(28) C_(32) C_(40) C_(48) C_(56) C_(64)
#undef C_
  }
  return cudaErrorInvalidValue;
}

}  // namespace pa
