// msv_tmem.cuh -- SSV/MSV filter with the emission table resident in Blackwell tensor memory.
//
// Why: the shared-memory SSV kernel (kernels.cuh, msv_kernel<Q>) is bound by the 128 B/clk/SM
// shared-memory crossbar (2 B of emission delta per DP cell) and, behind it, by the half-rate
// VIADDMNMX.S16x2 on the ALU pipe (ncu: smem wavefronts 82-86 %, ALU 63 %; profiles/).  On sm_100a
// two more resources are idle in that kernel and this one uses both:
//   * TMEM (256 KB/SM, 512 columns x 128 lanes x 32 bit) read with tcgen05.ld (SASS LDTM): measured
//     270 B/clk/SM (tools/probe_sm100.cu), twice the shared-memory crossbar, scoreboarded per
//     destination register like any load.  The striped table of one profile, (Kp+1) rows x Q
//     words per lane, is replicated into the four 32-lane quarters; a row fetch is ONE LDTM with a
//     warp-uniform column address (the residue), no bank conflicts possible.
//   * the FMA pipe: SSV values are small non-negative integers, exactly representable in fp16, so
//     u(i,k) = relu(u(i-1,k-1) + delta) is one HFMA2.RELU per two cells (a*1+c, relu fused);
//     non-negative fp16 bit patterns order like integers, so the running maximum stays a
//     VIMNMX3.S16x2 on the ALU pipe.  Both pipes issue concurrently (probe: 117 cells/clk/SM for the
//     8 HFMA2 + 4 VIMNMX3 mix against 63 for VIADDMNMX + VIMNMX3).
// Exactness: all values below 2048 are exact in fp16; a pair whose true SSV value reaches the
// candidate threshold Cp (<= 255) is re-run with the exact unsigned-byte MSV recurrence (also in
// fp16 lanes, values < 256), so rounding above 2048 can only happen to pairs that are already
// candidates.  Results are bit-identical to msv_kernel<Q> (tests/test_gpu_filters.py).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace pa {

// ---- tcgen05 wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t *slot, uint32_t ncols) {  // ncols: power of two >= 32
  if (threadIdx.x < 32) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_cta_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.ld 32x32b.xN: lane l of the warp receives N consecutive 32-bit columns of TMEM lane l
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld<1>(uint32_t taddr, uint32_t (&r)[1]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<2>(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
                 "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
                 "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
}
// tcgen05.wait::ld: the PTX model leaves the destination registers undefined until this point; the
// "+r" operands make every later use of the row depend on it.  ptxas turns it into scoreboard
// waits on exactly those registers, so younger loads stay in flight.
template <int N>
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[N]) {
  if constexpr (N == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0])::"memory");
  else if constexpr (N == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1])::"memory");
  else if constexpr (N == 4) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
  else {
#pragma unroll
    for (int o = 0; o < N; o += 8)
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(r[o]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])::"memory");
  }
}

// relu(a + c) on two fp16 lanes: HFMA2.RELU a, 1, c
__device__ __forceinline__ uint32_t h2_add_relu(uint32_t a, uint32_t c) {
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0x3c003c00u), "r"(c));
  return d;
}
__device__ __forceinline__ int h16_bits_to_int(int bits) { return __half2int_rn(__ushort_as_half((unsigned short)bits)); }
__device__ __forceinline__ uint32_t int_to_h16x2(int v) { return (uint32_t)__half_as_ushort(__int2half_rn(v)) * 0x10001u; }

__host__ __device__ constexpr int msv_tmem_lw(int Q) { return Q <= 1 ? 1 : Q <= 2 ? 2 : Q <= 4 ? 4 : Q <= 8 ? 8 : Q <= 16 ? 16 : 32; }
#define PA_TMEM_MAXQ 32
#define PA_TMEM_FULLQ 26   // widest all-TMEM row: 19 rows (DNA: Kp = 18 + the dead row) x 26 words + over-read <= 512 columns

// Geometry per Q (words per lane, Q = M/64 + 1):
//   Q <= 8    table needs <= 256 TMEM columns: two 512-thread CTAs per SM, 64 registers each
//   Q 9..16   all 512 columns: one 1024-thread CTA per SM, 64 registers
//   Q 17..32  hybrid: the first 16 words of every row come from TMEM, the other Q-16 from shared
//             memory (LDS.128, the striped layout of msv_kernel); one 512-thread CTA per SM
//   Q 17..26, FT ("full TMEM"): small alphabets (DNA: 19 rows) fit up to 26 words per row in the 512 columns, so the whole
//             row is ONE LDTM.x32 and no shared-memory part exists; one 512-thread CTA per SM, one row of prefetch (G = 1)
template <int Q, bool FT = false>
struct MsvTmemCfg {
  static constexpr int QT = (Q <= 16 || FT) ? Q : 16;
  static constexpr int QS = Q - QT;
  static constexpr int LW = msv_tmem_lw(QT);
  static constexpr int THREADS = Q <= 8 ? 512 : Q <= 16 ? 1024 : 512;
  static constexpr int MINB = Q <= 8 ? 2 : 1;
  static constexpr int G = (FT && Q > 16) ? 1 : 2;
};
__host__ __device__ constexpr int msv_tmem_qt(int Q, bool ft = false) { return (Q <= 16 || ft) ? Q : 16; }
// TMEM columns: (nrows + 1 dead row) * QT, plus the over-read of the last row's LDTM
__host__ __device__ constexpr int msv_tmem_cols_needed(int Q, int nrows, bool ft = false) {
  return (nrows + 1) * msv_tmem_qt(Q, ft) + (msv_tmem_lw(msv_tmem_qt(Q, ft)) - msv_tmem_qt(Q, ft));
}
// can rows of Q words live in TMEM entirely for an alphabet with `nrows` codes?
__host__ __device__ constexpr bool msv_tmem_full_ok(int Q, int nrows) {
  return Q > 16 && Q <= PA_TMEM_FULLQ && msv_tmem_cols_needed(Q, nrows, true) <= 512;
}
// words of the fp16x2 pool per profile: TMEM part [x][q < QT][lane], then the shared-memory part
// [x][striped QS*32] (rows 0..nrows-1; the dead row is synthesised on the device)
__host__ __device__ constexpr size_t msv_tmem_pool_words(int Q, int nrows) { return (size_t)nrows * Q * 32; }

// One SSV row for the whole warp: lane l owns cells (2l+hf)*Q + q (hf = half of word q).
template <int Q, int QT, int LW>
__device__ __forceinline__ void ssv_row_h2(uint32_t (&w)[Q], uint32_t &m0, uint32_t &m1, const uint32_t (&dt)[LW],
                                           const uint32_t *srow, int lane, int lanem1) {
  constexpr int QS = Q - QT;
  uint32_t ds[QS > 0 ? QS : 1];
  if constexpr (QS > 0) msv_load_row<QS>(srow, lane, ds);
  const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
  const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
#pragma unroll
  for (int q = Q - 1; q >= 1; q--) w[q] = h2_add_relu(w[q - 1], q < QT ? dt[q < QT ? q : 0] : ds[q >= QT ? q - QT : 0]);
  w[0] = h2_add_relu(carry, dt[0]);
#pragma unroll
  for (int q = 0; q + 1 < Q; q += 2) {
    if (q & 2) m1 = __vimax3_s16x2(m1, w[q], w[q + 1]);
    else m0 = __vimax3_s16x2(m0, w[q], w[q + 1]);
  }
  if (Q & 1) m1 = __vmaxs2(m1, w[Q - 1]);
}

// relu(d * s + c) on two fp16 lanes: HFMA2.RELU d, s, c
__device__ __forceinline__ uint32_t h2_fma_relu(uint32_t d, uint32_t s, uint32_t c) {
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(d), "r"(s), "r"(c));
  return r;
}

// "Sticky" SSV row: no running maximum at all.  Cell values are kept as n * 32 with n = u * m, m = ceil(2048 / Cp)
// (Cp = the pair's candidate threshold, 2..255): n < 2048 is an exact fp16 integer multiple of 32, and the first
// cell whose true value reaches u >= ceil(2048 / m) (<= Cp) produces n * 32 >= 65536, which rounds to +inf.  inf is
// absorbing under fma with a finite product, so it rides down the diagonal (through dead cells, around the striped
// model) until the last row, where one look at the registers tells whether ANY cell of the pair reached the
// threshold.  The scale s = m * 32 is the second fma operand (the table keeps the plain deltas), so the row is
// 16 HFMA2 + the carry for 1,024 cells instead of 16 HFMA2 + 8 VIMNMX3.  A pair flagged this way is a superset of
// the exact candidate set (threshold ceil(2048 / m) <= Cp); candidates are re-run with the exact recurrence.
template <int Q, int QT, int LW>
__device__ __forceinline__ void ssv_row_sticky(uint32_t (&w)[Q], const uint32_t sc, const uint32_t (&dt)[LW], const uint32_t *srow,
                                               int lane, int lanem1) {
  constexpr int QS = Q - QT;
  uint32_t ds[QS > 0 ? QS : 1];
  if constexpr (QS > 0) msv_load_row<QS>(srow, lane, ds);
  const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
  const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
#pragma unroll
  for (int q = Q - 1; q >= 1; q--) w[q] = h2_fma_relu(q < QT ? dt[q < QT ? q : 0] : ds[q >= QT ? q - QT : 0], sc, w[q - 1]);
  w[0] = h2_fma_relu(dt[0], sc, carry);
}
// max over the row registers, as fp16 bit patterns (all values >= 0; +inf = 0x7c00, NaN = 0x7fff)
template <int Q>
__device__ __forceinline__ uint32_t row_max_bits(const uint32_t (&w)[Q]) {
  uint32_t m = w[0];
#pragma unroll
  for (int q = 1; q + 1 < Q; q += 2) m = __vimax3_s16x2(m, w[q], w[q + 1]);
  if ((Q & 1) == 0) m = __vmaxs2(m, w[Q - 1]);
  return m;
}

// G rows per prefetch group (double-buffered): the LDTMs of group g+1 are in flight while group g
// is computed.  The residues of the next 32-row block (of this target or of the warp's next target)
// and the next target's metadata are fetched one block ahead, so no global-memory latency sits on
// the per-target critical path.
template <int Q, int G, bool FT = false>
__global__ void __launch_bounds__(MsvTmemCfg<Q, FT>::THREADS, MsvTmemCfg<Q, FT>::MINB) msv_tmem_kernel(MsvArgs a, int ncols) {
  using Cfg = MsvTmemCfg<Q, FT>;
  constexpr int QT = Cfg::QT, QS = Cfg::QS, LW = Cfg::LW;
  constexpr int SROW = QS * 32;                       // words per shared-memory row
  extern __shared__ __align__(16) uint32_t s_tab[];  // (nrows + 1) * SROW words (hybrid only)
  __shared__ uint32_t s_slot;
  __shared__ unsigned long long s_item;
  __shared__ unsigned int s_next;
  // the shuffle tells the compiler that the warp index (and everything derived from it: target index,
  // length, loop bounds) is warp-uniform, so shuffles inside those loops need no divergence guards
  const int lane = threadIdx.x & 31, wib = __shfl_sync(PA_FULL, (int)(threadIdx.x >> 5), 0), wpb = blockDim.x >> 5;
  const uint32_t tbase = tmem_alloc(&s_slot, (uint32_t)ncols);
  const uint32_t lanebase = tbase + ((uint32_t)((wib & 3) * 32) << 16);
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)a.nlist * ntiles;
  int cur_h = -1;
  unsigned long long ncand = 0;
  const int lanem1 = (lane + 31) & 31;
  // work items (profile-major) are handed out through an atomic counter: CTAs finish within one item of
  // each other, and neighbouring items share the profile, so the TMEM table is refilled rarely.  Inside an
  // item the warps claim targets one at a time from a shared-memory counter, so a warp that has to re-run a
  // candidate with the exact recurrence does not keep the others waiting at the item's barrier.
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { s_item = atomicAdd(a.work, 1ull); s_next = (unsigned int)wpb; }
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= nitems) break;
    uint32_t slot, tile;
    msv_item_decode(item, (uint32_t)a.nlist, ntiles, a.super_tiles, slot, tile);
    const int h = a.hmm_list[slot];
    const DevHmm &hm = a.hmms[h];
    const int nrows = hm.nrows;
    if (h != cur_h) {
      tmem_cta_sync();  // every warp is done with the previous table
      const uint32_t *src = a.tab + hm.msvh_off;
      const uint32_t dead = int_to_h16x2(PA_MSV_DEAD);
      if (wib < 4) {
        for (int x = 0; x < nrows; x++)
#pragma unroll
          for (int q = 0; q < QT; q++) tmem_st1(lanebase + x * QT + q, __ldg(src + (x * QT + q) * 32 + lane));
#pragma unroll
        for (int q = 0; q < LW; q++) tmem_st1(lanebase + nrows * QT + q, dead);
        tmem_wait_st();
      }
      if constexpr (QS > 0) {
        const uint32_t *ssrc = src + (size_t)nrows * QT * 32;
        for (int i = threadIdx.x; i < nrows * SROW; i += blockDim.x) s_tab[i] = __ldg(ssrc + i);
        for (int i = threadIdx.x; i < SROW; i += blockDim.x) s_tab[nrows * SROW + i] = dead;
      }
      tmem_cta_sync();
      cur_h = h;
    }
    const int bias = hm.bias, base = hm.base, tec = hm.tec, tbm = hm.tbm;
    const uint32_t t0 = tile * a.tile, t1 = min(a.tg.n, (tile + 1) * a.tile);
    uint32_t t = t0 + wib, tn;
    int nxL = 0, nxli = 0;
    uint32_t nxoff = 0;
    if (t < t1) { nxL = (int)a.tg.len[t]; nxoff = a.tg.off[t]; nxli = a.tg.lidx[t]; }
    int pre_x = 0;          // this lane's residue of the next block to be processed
    bool have_pre = false;  // warp-uniform
    for (; t < t1; t = tn) {
      const int L = nxL, li = nxli;
      const uint8_t *p = a.tg.dsq + nxoff;
      {  // claim this warp's next target and fetch its metadata one target ahead
        unsigned int c = 0;
        if (lane == 0) c = atomicAdd(&s_next, 1u);
        tn = t0 + __shfl_sync(PA_FULL, c, 0);
      }
      if (tn < t1) { nxL = (int)a.tg.len[tn]; nxoff = a.tg.off[tn]; nxli = a.tg.lidx[tn]; }
      else nxL = 0;
      if (L == 0) { have_pre = false; continue; }
      const int tjb = a.tjb_li[li];
      const int T = a.thrJ[(size_t)h * a.nL + li];
      const int tjbm = (tjb + tbm) & 0xff;
      const int xB0 = max(base - tjbm, 0);
      const int C = min(255 - bias, min(base + 1 + tec, T + tec));
      const int Cp = a.all_pairs ? 0 : C - xB0;
      uint32_t w[Q];
      if (Cp > 0) {
        // ---- SSV pass: u(i,k) = relu(u(i-1,k-1) + delta(x_i,k)), running max over all cells
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = 0;
        const uint32_t sc = (uint32_t)a.scJ[(size_t)h * a.nL + li] * 0x10001u;
        if (!have_pre) pre_x = lane < L ? (int)p[lane] : nrows;
        for (int i0 = 0; i0 < L; i0 += 32) {
          const int nr = min(32, L - i0);
          // lane r holds row i0+r: its residue (rows past the end: the dead row)
          const int myx = pre_x;
          if (i0 + 32 < L) pre_x = i0 + 32 + lane < L ? (int)p[i0 + 32 + lane] : nrows;
          else { pre_x = lane < nxL ? (int)a.tg.dsq[nxoff + lane] : nrows; have_pre = true; }
          const uint32_t myaddr = lanebase + (uint32_t)(myx * QT);
          uint32_t buf[2][G][LW];
#pragma unroll
          for (int g = 0; g < G; g++) tmem_ld<LW>(__shfl_sync(PA_FULL, myaddr, g), buf[0][g]);
          // the 32-row block is fully unrolled (row indices are immediates, no register rotation copies);
          // the prefetch of the next group is unconditional: past the end it fetches the dead row
#pragma unroll
          for (int r = 0; r < 32; r += G) {
            if (r >= nr) break;
            if (r + G < 32) {
#pragma unroll
              for (int g = 0; g < G; g++) tmem_ld<LW>(__shfl_sync(PA_FULL, myaddr, r + G + g), buf[((r / G) & 1) ^ 1][g]);
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
              const uint32_t *srow = s_tab;
              if constexpr (QS > 0) srow = s_tab + __shfl_sync(PA_FULL, myx, r + g) * SROW;
              tmem_wait_ld<LW>(buf[(r / G) & 1][g]);
              ssv_row_sticky<Q, QT, LW>(w, sc, buf[(r / G) & 1][g], srow, lane, lanem1);
            }
          }
        }
        const uint32_t mx = row_max_bits<Q>(w);
        if (!__any_sync(PA_FULL, (mx & 0xffffu) >= 0x7c00u || (mx >> 16) >= 0x7c00u)) continue;
      } else
        have_pre = false;
      ncand++;
      // ---- exact MSV recurrence (J state, overflow) for candidates; values < 256, exact in fp16
      int xJ = 0, xB = xB0;
      bool overflow = false;
#pragma unroll
      for (int q = 0; q < Q; q++) w[q] = 0;
      for (int i = 0; i < L; i++) {
        const int x = p[i];
        uint32_t d[LW], ds[QS > 0 ? QS : 1];
        tmem_ld<LW>(lanebase + (uint32_t)(x * QT), d);
        if constexpr (QS > 0) msv_load_row<QS>(s_tab + x * SROW, lane, ds);
        const uint32_t xBv = int_to_h16x2(xB);
        const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
        const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
        tmem_wait_ld<LW>(d);
        uint32_t rm = 0;
#pragma unroll
        for (int q = Q - 1; q >= 1; q--) {
          w[q] = h2_add_relu(__vmaxs2(w[q - 1], xBv), q < QT ? d[q < QT ? q : 0] : ds[q >= QT ? q - QT : 0]);
          rm = __vmaxs2(rm, w[q]);
        }
        w[0] = h2_add_relu(__vmaxs2(carry, xBv), d[0]);
        rm = __vmaxs2(rm, w[0]);
        int xE = h16_bits_to_int(warp_max16x2(rm));
        if (xE >= 255 - bias) { overflow = true; break; }
        xE = max(xE - tec, 0);
        xJ = max(xJ, xE);
        xB = max(max(base, xJ) - tjbm, 0);
      }
      if (overflow) xJ = -1;
      if (a.all_pairs || overflow || xJ >= T) {
        if (lane == 0) {
          unsigned long long idx = atomicAdd(a.out_count, 1ull);
          if (idx < a.out_cap) {
            Pass1 r;
            r.t = t; r.h = (uint16_t)h; r.xJ = (int16_t)xJ;
            a.out[idx] = r;
          }
        }
      }
    }
  }
  if (lane == 0 && ncand) atomicAdd(a.n_ssv_cand, ncand);
  tmem_dealloc(tbase, (uint32_t)ncols);
}

// ---------------------------------------------------------------------------------------------
// Packed profiles.  The per-row cost of the SSV loop has a fixed part (address shuffle, R2UR, LDTM,
// carry shuffle, loop) that is amortised over Q words, so narrow profiles run at 55-69 cells/clk/SM
// while Q = 16 reaches 89.  A *pack* concatenates several profiles into one 1024-cell striped model
// with Q = 16: profile p owns the half-lanes [s, s + n), n = M/16 + 1, so that (i) every half-lane
// (16 consecutive cells held by one 16-bit half of a lane) belongs to one profile, (ii) the last cell of
// every profile is dead padding (delta -20000), which is exactly the u(i-1, 0) = 0 boundary of the next
// profile's first column, (iii) padding is at most 16 cells per profile instead of 64.  The two 16-bit
// halves of the running maximum stay separate in VIMNMX3.S16x2, so the per-profile maximum is a masked
// warp reduction after the last row.  Candidates of a member are re-run with the exact MSV recurrence
// on the pack's table; only the member's own half-lanes enter its row maximum.
struct DevPack {
  unsigned long long msvh_off;  // fp16x2 table [x][q < 16][lane]
  int nrows, nmem, mem_off;     // members: d_packmem[mem_off + 3*i] = {profile index, first half-lane, half-lanes}
};
struct MsvPackArgs {
  MsvArgs a;  // hmm_list/nlist unused; tab = fp16 pool
  const DevPack *packs;
  const int *packmem;
  int npacks;
};

// QP = words per lane of the packed model = cells per half-lane: 16 (any alphabet, 1024 threads) or 26 (alphabets whose table
// fits TMEM at that width, i.e. DNA: 1,664-cell packs, one LDTM.x32 per row, 512 threads with one row of prefetch)
template <int QP, int G>
__global__ void __launch_bounds__(QP <= 16 ? 1024 : 512, 1) msv_pack_kernel(MsvPackArgs pa_, int ncols) {
  constexpr int Q = QP, QT = QP, LW = msv_tmem_lw(QP);
  const MsvArgs &a = pa_.a;
  __shared__ uint32_t s_slot;
  __shared__ unsigned long long s_item;
  __shared__ unsigned int s_next;
  const int lane = threadIdx.x & 31, wib = __shfl_sync(PA_FULL, (int)(threadIdx.x >> 5), 0), wpb = blockDim.x >> 5;
  const uint32_t tbase = tmem_alloc(&s_slot, (uint32_t)ncols);
  const uint32_t lanebase = tbase + ((uint32_t)((wib & 3) * 32) << 16);
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)pa_.npacks * ntiles;
  int cur_pk = -1;
  unsigned long long ncand = 0;
  const int lanem1 = (lane + 31) & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { s_item = atomicAdd(a.work, 1ull); s_next = (unsigned int)wpb; }   // warp w starts with target w of the item
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= nitems) break;
    uint32_t slot, tile;
    msv_item_decode(item, (uint32_t)pa_.npacks, ntiles, a.super_tiles, slot, tile);
    const int pk = (int)slot;
    const DevPack pack = pa_.packs[pk];
    const int nrows = pack.nrows;
    const int *mem = pa_.packmem + pack.mem_off;
    // owners of this lane's two half-lanes (-1: none), after the member triples
    const int own_lo = mem[3 * pack.nmem + 2 * lane], own_hi = mem[3 * pack.nmem + 2 * lane + 1];
    if (pk != cur_pk) {
      tmem_cta_sync();
      if (wib < 4) {
        const uint32_t *src = a.tab + pack.msvh_off;
        for (int x = 0; x < nrows; x++)
#pragma unroll
          for (int q = 0; q < QT; q++) tmem_st1(lanebase + x * QT + q, __ldg(src + (x * QT + q) * 32 + lane));
        const uint32_t dead = int_to_h16x2(PA_MSV_DEAD);
#pragma unroll
        for (int q = 0; q < LW; q++) tmem_st1(lanebase + nrows * QT + q, dead);
        tmem_wait_st();
      }
      tmem_cta_sync();
      cur_pk = pk;
    }
    const uint32_t t0 = tile * a.tile, t1 = min(a.tg.n, (tile + 1) * a.tile);
    uint32_t t = t0 + wib, tn;
    int nxL = 0, nxli = 0;
    uint32_t nxoff = 0;
    if (t < t1) { nxL = (int)a.tg.len[t]; nxoff = a.tg.off[t]; nxli = a.tg.lidx[t]; }
    int pre_x = 0;
    bool have_pre = false;
    for (; t < t1; t = tn) {
      const int L = nxL, li = nxli;
      const uint8_t *p = a.tg.dsq + nxoff;
      {  // targets of the item are claimed one at a time (shared-memory counter): a warp busy with an exact re-run
         // does not hold up the item; the next target's metadata and first residues are fetched one target ahead
        unsigned int c = 0;
        if (lane == 0) c = atomicAdd(&s_next, 1u);
        tn = t0 + __shfl_sync(PA_FULL, c, 0);
      }
      if (tn < t1) { nxL = (int)a.tg.len[tn]; nxoff = a.tg.off[tn]; nxli = a.tg.lidx[tn]; }
      else nxL = 0;
      if (L == 0) { have_pre = false; continue; }
      // ---- SSV pass over the whole pack
      uint32_t w[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) w[q] = 0;
      // per half-lane scale of its owner for this length (0xffff: the owner takes every pair, Cp <= 0)
      const uint32_t sc_lo = own_lo >= 0 ? (uint32_t)a.scJ[(size_t)own_lo * a.nL + li] : 0u;
      const uint32_t sc_hi = own_hi >= 0 ? (uint32_t)a.scJ[(size_t)own_hi * a.nL + li] : 0u;
      const uint32_t sc = sc_lo | (sc_hi << 16);
      const bool always = __any_sync(PA_FULL, sc_lo == 0xffffu || sc_hi == 0xffffu);
      if (!have_pre) pre_x = lane < L ? (int)p[lane] : nrows;
      for (int i0 = 0; i0 < L; i0 += 32) {
        const int nr = min(32, L - i0);
        const int myx = pre_x;
        if (i0 + 32 < L) pre_x = i0 + 32 + lane < L ? (int)p[i0 + 32 + lane] : nrows;
        else { pre_x = lane < nxL ? (int)a.tg.dsq[nxoff + lane] : nrows; have_pre = true; }
        const uint32_t myaddr = lanebase + (uint32_t)(myx * QT);
        uint32_t buf[2][G][LW];
#pragma unroll
        for (int g = 0; g < G; g++) tmem_ld<LW>(__shfl_sync(PA_FULL, myaddr, g), buf[0][g]);
#pragma unroll
        for (int r = 0; r < 32; r += G) {
          if (r >= nr) break;
          if (r + G < 32) {
#pragma unroll
            for (int g = 0; g < G; g++) tmem_ld<LW>(__shfl_sync(PA_FULL, myaddr, r + G + g), buf[((r / G) & 1) ^ 1][g]);
          }
#pragma unroll
          for (int g = 0; g < G; g++) {
            tmem_wait_ld<LW>(buf[(r / G) & 1][g]);
            ssv_row_sticky<Q, QT, LW>(w, sc, buf[(r / G) & 1][g], nullptr, lane, lanem1);
          }
        }
      }
      // half-lanes that hold an absorbed +inf after the last row: bit 2*lane (+1)
      const uint32_t mx = row_max_bits<Q>(w);
      const uint32_t even = __ballot_sync(PA_FULL, (mx & 0xffffu) >= 0x7c00u && sc_lo != 0xffffu);
      const uint32_t odd = __ballot_sync(PA_FULL, (mx >> 16) >= 0x7c00u && sc_hi != 0xffffu);
      if ((even | odd) == 0u && !always) continue;
      unsigned long long hot = 0;
      for (int b = 0; b < 32; b++) hot |= ((unsigned long long)((even >> b) & 1u) << (2 * b)) | ((unsigned long long)((odd >> b) & 1u) << (2 * b + 1));
      // an inf seen in cell c was born at most L-1 cells up the diagonal: a member owning half-lanes [hs, hs+hn) is a
      // candidate iff some hot half-lane lies in [hs, hs + hn + ceil((L-1)/QP)] (cyclic over the 64 half-lanes)
      const int reach = (L + QP - 2) / QP;
      const int tjb = a.tjb_li[li];
      // ---- per member: candidate test, exact MSV re-run
      for (int mi = 0; mi < pack.nmem; mi++) {
        const int h = mem[3 * mi], hs = mem[3 * mi + 1], hn = mem[3 * mi + 2];
        const DevHmm &hm = a.hmms[h];
        const int bias = hm.bias, base = hm.base, tec = hm.tec, tbm = hm.tbm;
        const int T = a.thrJ[(size_t)h * a.nL + li];
        const int tjbm = (tjb + tbm) & 0xff;
        const int xB0 = max(base - tjbm, 0);
        const int C = min(255 - bias, min(base + 1 + tec, T + tec));
        const int Cp = C - xB0;
        const bool lo_in = (unsigned)(2 * lane - hs) < (unsigned)hn, hi_in = (unsigned)(2 * lane + 1 - hs) < (unsigned)hn;
        if (Cp > 0) {
          const int cnt = hn + reach;
          const unsigned long long win0 = cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull);
          const unsigned long long win = hs ? ((win0 << hs) | (win0 >> (64 - hs))) : win0;
          if ((hot & win) == 0ull) continue;
        }
        ncand++;
        const uint32_t keep = (lo_in ? 0x0000ffffu : 0u) | (hi_in ? 0xffff0000u : 0u);
        int xJ = 0, xB = xB0;
        bool overflow = false;
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = 0;
        for (int i = 0; i < L; i++) {
          const int x = p[i];
          uint32_t d[LW];
          tmem_ld<LW>(lanebase + (uint32_t)(x * QT), d);
          const uint32_t xBv = int_to_h16x2(xB);
          const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
          const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
          tmem_wait_ld<LW>(d);
          uint32_t rm = 0;
#pragma unroll
          for (int q = Q - 1; q >= 1; q--) {
            w[q] = h2_add_relu(__vmaxs2(w[q - 1], xBv), d[q]);
            rm = __vmaxs2(rm, w[q]);
          }
          w[0] = h2_add_relu(__vmaxs2(carry, xBv), d[0]);
          rm = __vmaxs2(rm, w[0]) & keep;   // only this member's half-lanes
          int xE = h16_bits_to_int(warp_max16x2(rm));
          if (xE >= 255 - bias) { overflow = true; break; }
          xE = max(xE - tec, 0);
          xJ = max(xJ, xE);
          xB = max(max(base, xJ) - tjbm, 0);
        }
        if (overflow) xJ = -1;
        if (overflow || xJ >= T) {
          if (lane == 0) {
            unsigned long long idx = atomicAdd(a.out_count, 1ull);
            if (idx < a.out_cap) {
              Pass1 r;
              r.t = t; r.h = (uint16_t)h; r.xJ = (int16_t)xJ;
              a.out[idx] = r;
            }
          }
        }
      }
    }
  }
  if (lane == 0 && ncand) atomicAdd(a.n_ssv_cand, ncand);
  tmem_dealloc(tbase, (uint32_t)ncols);
}

}  // namespace pa
