// vit16.cuh -- packed 16-bit upper-bound pass in front of the exact Viterbi filter.
//
// The Viterbi filter only decides P(score) <= F2, and ~98 % of the pairs that reach it fail.  vit_kernel
// evaluates HMMER's saturating signed-16 recurrences exactly, one cell per s32 lane-op.  This kernel evaluates the same
// recurrences two cells per instruction (VIADDMNMX.S16x2 / VIMNMX.S16x2) under rules that can only RAISE a value:
//   * every stored value is floored at -16384 and every table entry is clamped to >= -16384, so a + b >= -32768: the
//     16-bit add never wraps at the bottom (HMMER's floor is -32768; values that low belong to dead paths);
//   * the row maximum is tested against the pair's pass threshold after every row; reaching it ends the pass with the
//     verdict "candidate", so no value ever exceeds threshold + the largest emission and the add never wraps at the top.
// Max-plus recurrences are monotone in their inputs, hence the result is an UPPER BOUND of the exact xC: a pair whose
// bound stays below the pass threshold fails the exact filter as well and is dropped here; the others (a few per cent)
// are flagged and re-run by the exact vit_kernel, whose decision is final.  The passing set is therefore identical.
// Layout: 64 half-lanes; half-lane hl = 2*lane + hf owns cells k = hl*BH + j + 1 (j < BH), word j of a lane packs its two
// half-lanes, so the (i-1, k-1) dependency is word j-1 of the same register file and one shuffle + PRMT per row at j = 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace pa {

#define PA_V16_FLOOR (-16384)

struct Vit16Args {
  const DevHmm *hmms;
  const uint32_t *tab;     // v16 pool: per profile trA(uint4) | trI(uint2) | trD(uint2) | ps | em[Kp], each [j][lane]
  Targets tg;
  LenTabs lt;
  const Pass2 *in;         // sorted by profile rank; this launch handles [begin, end)
  unsigned long long begin, end;
  uint8_t *need;           // per input index: 1 = the exact kernel must process this pair
  unsigned long long *n_need;
};

__device__ __forceinline__ uint32_t rep16(int v) { return (uint32_t)(v & 0xffff) * 0x10001u; }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)(short)(w >> 16); }
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffff); }

// returns true if the upper bound of xC reaches Tint
template <int BH>
__device__ __forceinline__ bool vit16_pair(const DevHmm &hm, const uint32_t *tab, const uint8_t *p, int L, int xw_move, int Tint, int lane) {
  constexpr int W = BH * 32;
  const uint4 *trA = reinterpret_cast<const uint4 *>(tab + hm.v16_off);
  const uint2 *trI = reinterpret_cast<const uint2 *>(trA + W);
  const uint2 *trD = trI + W;
  const uint32_t *ps = reinterpret_cast<const uint32_t *>(trD + W);
  const uint32_t *em = ps + W;
  const uint32_t FLv = rep16(PA_V16_FLOOR);
  uint32_t Mx[BH], Ix[BH], Dx[BH];
#pragma unroll
  for (int j = 0; j < BH; j++) { Mx[j] = FLv; Ix[j] = FLv; Dx[j] = FLv; }
  const int xN = hm.base_w, xw_E = hm.xw_E;
  int xB = max(xN + xw_move, PA_V16_FLOOR), xJ = PA_V16_FLOOR;
  const uint32_t pslast = __ldg(ps + (BH - 1) * 32 + lane);
  const int lanem1 = (lane + 31) & 31;
  for (int i = 0; i < L; i++) {
    const uint32_t *emx = em + (size_t)p[i] * W;
    const uint32_t xBv = rep16(xB);
    // previous-row values of the cell left of word 0: lo half <- hi half of lane-1, hi half <- own lo half
    uint32_t mL = __byte_perm(__shfl_sync(PA_FULL, Mx[BH - 1], lanem1), Mx[BH - 1], 0x5432);
    uint32_t iL = __byte_perm(__shfl_sync(PA_FULL, Ix[BH - 1], lanem1), Ix[BH - 1], 0x5432);
    uint32_t dL = __byte_perm(__shfl_sync(PA_FULL, Dx[BH - 1], lanem1), Dx[BH - 1], 0x5432);
    if (lane == 0) {  // nothing left of cell 1
      mL = (mL & 0xffff0000u) | (FLv & 0xffffu);
      iL = (iL & 0xffff0000u) | (FLv & 0xffffu);
      dL = (dL & 0xffff0000u) | (FLv & 0xffffu);
    }
    uint32_t rowmax = FLv;
#pragma unroll
    for (int j = BH - 1; j >= 0; j--) {
      const int c = j * 32 + lane;
      const uint4 A = __ldg(trA + c);
      const uint2 Iq = __ldg(trI + c);
      const uint32_t e = __ldg(emx + c);
      const uint32_t pm = j ? Mx[j > 0 ? j - 1 : 0] : mL, pi = j ? Ix[j > 0 ? j - 1 : 0] : iL, pd = j ? Dx[j > 0 ? j - 1 : 0] : dL;
      uint32_t sv = __viaddmax_s16x2(xBv, A.x, FLv);
      sv = __viaddmax_s16x2(pm, A.y, sv);
      sv = __viaddmax_s16x2(pi, A.z, sv);
      sv = __viaddmax_s16x2(pd, A.w, sv);
      sv = __viaddmax_s16x2(sv, e, FLv);
      uint32_t iv = __viaddmax_s16x2(Mx[j], Iq.x, FLv);
      iv = __viaddmax_s16x2(Ix[j], Iq.y, iv);
      Mx[j] = sv;
      Ix[j] = iv;
      rowmax = __vmaxs2(rowmax, sv);
    }
    const int xE = __reduce_max_sync(PA_FULL, max(lo16(rowmax), hi16(rowmax)));
    if (xE + xw_E >= Tint) return true;
    // ---- D: serial inside each half-lane, then a max-plus scan over the 64 half-lanes
    uint32_t mcL = __byte_perm(__shfl_sync(PA_FULL, Mx[BH - 1], lanem1), Mx[BH - 1], 0x5432);
    if (lane == 0) mcL = (mcL & 0xffff0000u) | (FLv & 0xffffu);
    uint32_t aloc = FLv;
#pragma unroll
    for (int j = 0; j < BH; j++) {
      const uint2 Dq = __ldg(trD + j * 32 + lane);
      const uint32_t mleft = j ? Mx[j > 0 ? j - 1 : 0] : mcL;
      const uint32_t t = __viaddmax_s16x2(mleft, Dq.x, FLv);
      aloc = __viaddmax_s16x2(aloc, Dq.y, t);
      Dx[j] = aloc;
    }
    // element of half-lane: (A = D of its last cell when nothing enters from the left, S = sum of its DD, floored)
    const int Alo = lo16(aloc), Ahi0 = hi16(aloc), Slo = lo16(pslast), Shi0 = hi16(pslast);
    // inclusive prefix over (lo, hi) of this lane, then Hillis-Steele across lanes on the lane totals
    int LA = max(Ahi0, max(Alo + Shi0, PA_V16_FLOOR)), LS = max(Slo + Shi0, PA_V16_FLOOR);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a2 = __shfl_up_sync(PA_FULL, LA, o), s2 = __shfl_up_sync(PA_FULL, LS, o);
      if (lane >= o) {
        LA = max(LA, max(a2 + LS, PA_V16_FLOOR));
        LS = max(LS + s2, PA_V16_FLOOR);
      }
    }
    int din_lo = __shfl_up_sync(PA_FULL, LA, 1);           // D leaving half-lane 2*lane - 1
    if (lane == 0) din_lo = PA_V16_FLOOR;
    const int din_hi = max(Alo, max(din_lo + Slo, PA_V16_FLOOR));  // D leaving half-lane 2*lane
    const uint32_t din = (uint32_t)(din_lo & 0xffff) | ((uint32_t)(din_hi & 0xffff) << 16);
#pragma unroll
    for (int j = 0; j < BH; j++) Dx[j] = __viaddmax_s16x2(din, __ldg(ps + j * 32 + lane), Dx[j]);
    xJ = max(xJ, max(xE + xw_E, PA_V16_FLOOR));
    xB = max(max(xJ + xw_move, xN + xw_move), PA_V16_FLOOR);
  }
  return false;   // xC = max over rows of xE + xw_E stayed below Tint
}

template <int BH>
__global__ void __launch_bounds__(128, (BH <= 4 ? 6 : BH <= 8 ? 4 : 3)) vit16_kernel(Vit16Args a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  unsigned long long nneed = 0;
  for (unsigned long long i = a.begin + (unsigned long long)blockIdx.x * wpb + wib; i < a.end; i += (unsigned long long)gridDim.x * wpb) {
    const Pass2 r = a.in[i];
    const DevHmm &hm = a.hmms[r.h];
    bool need = true;
    if (r.flags & 1) {   // otherwise P <= F2 already: the exact kernel forwards it without running Viterbi
      const int L = (int)a.tg.len[r.t];
      const int xw_move = a.lt.xw_move[a.tg.lidx[r.t]];
      // smallest xC that can pass: seq = ((xC + xw_move - base_w) / scale_w - 3 - filtersc) / ln2 >= thrF2v; two units of slack
      // cover the float rounding of the exact kernel's own evaluation
      const float need_sc = (float)((double)hm.thrF2v * PA_LN2) + r.filtersc + 3.0f;
      const float tf = need_sc * hm.scale_w + (float)hm.base_w - (float)xw_move;
      // thresholds outside the 16-bit working range (F2 >= 1 gives -inf, absurdly strict ones exceed 32767 - guard): exact kernel
      if (tf > (float)(PA_V16_FLOOR + 8) && tf < 32000.0f) {
        const int Tint = (int)floorf(tf) - 2;
        if (Tint + hm.v16_guard < 32767) need = vit16_pair<BH>(hm, a.tab, a.tg.dsq + a.tg.off[r.t], L, xw_move, Tint, lane);
      }
    }
    if (lane == 0) a.need[i] = need ? 1 : 0;
    nneed += need ? 1 : 0;
  }
  if (lane == 0 && nneed) atomicAdd(a.n_need, nneed);
}

}  // namespace pa
