// wide.cuh -- profiles wider than 2,047 columns (nucleotide / codon-mode references of ordinary genes: the
// reference's own fixture /root/reference/tests/ref/OG0003977.fa is 2,049-2,052 nt; hmmsearch, lib/aln.py:664-667,
// has no width limit).  The register-resident kernels of kernels.cuh / msv_tmem.cuh / vit16.cuh hold one striped
// model per warp, which caps them at 32 words (MSV), 32 packed words (vit16) or 64 cells (vit, fwd) per lane.
// Here the model is spread over more than one warp or swept tile by tile:
//
//  msv_wide_kernel<Q,G>   sticky-inf SSV pass, MODEL-TILED: the profile is cut into nT tiles of 64*Q cells
//                         (Q = 17..32, the hybrid TMEM + shared-memory geometry of msv_tmem_kernel).  A CTA keeps
//                         one tile's table resident and sweeps all rows of all its targets through it; the only
//                         thing that crosses a tile boundary in u(i,k) = relu(u(i-1,k-1) + delta) is the last
//                         cell of the previous row, one fp16 per (target, row), parked in an L2-resident scratch
//                         row.  +inf is absorbing, so a threshold crossing in any tile is still visible at the end
//                         (in the registers after the last row, or leaving through the right edge of the last
//                         tile).  Flagged pairs go to a candidate list.
//  msv_exact_wide_kernel  exact unsigned-byte MSV recurrence (J state, overflow) for the candidates, any width:
//                         striped state in shared memory, s16x2 DPX lanes, table rows from L2.
//  vit16_wide_kernel<BH>  the packed 16-bit upper-bound pass with the model spread over the NW warps of a CTA
//  vit_wide_kernel<BT>    the exact s32 Viterbi filter, same decomposition.  Cross-warp traffic per row: the row
//                         maximum, the word left of each warp's first cell, and the warp totals of the D->D
//                         max-plus scan -- two CTA barriers per row.  Max-plus is order-independent, so the results
//                         equal the single-warp kernels' bit for bit.
//  fwd_wide_kernel        Forward filter with the canonical 32-lane blocked float order and a
//                         run-time block size: one warp per CTA, rows in shared memory, table reads batched.
#pragma once
#include "kernels.cuh"
#include "msv_tmem.cuh"
#include "vit16.cuh"

namespace pa {

#define PA_MAX_M 8192   // widest profile: 8 warps x 32 lanes x 32 cells of vit_wide_kernel

// ---------------------------------------------------------------------------------------------
// MSV, model-tiled sticky-inf SSV pass
// ---------------------------------------------------------------------------------------------
struct MsvWideArgs {
  MsvArgs a;                 // hmm_list: wide profiles whose tiles have Q words per lane; tab: fp16 pool
  uint16_t *bnd;             // gridDim.x regions of a.tile rows of Lpad halfwords: bnd[row i] = value entering cell 0 at row i
  uint32_t Lpad;
  Pass1 *cand;               // flagged pairs {t, h}
  unsigned long long *cand_count;
  unsigned long long cand_cap;
};

template <int Q, int G, bool FT = false>
static __global__ void __launch_bounds__(512, 1) msv_wide_kernel(MsvWideArgs wa, int ncols) {
  static_assert(Q >= 17 && Q <= 32 && (!FT || Q <= PA_TMEM_FULLQ), "tiles use the hybrid TMEM + shared-memory geometry, or all-TMEM rows (FT)");
  using Cfg = MsvTmemCfg<Q, FT>;
  constexpr int QT = Cfg::QT, QS = Cfg::QS, LW = Cfg::LW, SROW = QS * 32;
  const MsvArgs &a = wa.a;
  extern __shared__ __align__(16) uint32_t s_tab[];  // (nrows + 1) * SROW words
  __shared__ uint32_t s_slot;
  __shared__ unsigned long long s_item;
  const int lane = threadIdx.x & 31, wib = __shfl_sync(PA_FULL, (int)(threadIdx.x >> 5), 0), wpb = blockDim.x >> 5;
  const uint32_t tbase = tmem_alloc(&s_slot, (uint32_t)ncols);
  const uint32_t lanebase = tbase + ((uint32_t)((wib & 3) * 32) << 16);
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)a.nlist * ntiles;
  unsigned long long ncand = 0;
  const int lanem1 = (lane + 31) & 31;
  uint16_t *bnd_cta = wa.bnd + (size_t)blockIdx.x * a.tile * wa.Lpad;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(a.work, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= nitems) break;
    uint32_t slot, tile;
    msv_item_decode(item, (uint32_t)a.nlist, ntiles, a.super_tiles, slot, tile);
    const int h = a.hmm_list[slot];
    const DevHmm &hm = a.hmms[h];
    const int nrows = hm.nrows, nT = hm.nT;
    const int bias = hm.bias, base = hm.base, tec = hm.tec, tbm = hm.tbm;
    const uint32_t t0 = tile * a.tile, t1 = min(a.tg.n, (tile + 1) * a.tile);
    uint32_t flags = 0;  // bit j: the j-th target of this warp in the tile is a candidate (a.tile <= 32 * wpb)
    for (int m = 0; m < nT; m++) {
      tmem_cta_sync();  // every warp is done with the previous table
      {
        const uint32_t *src = a.tab + hm.msvh_off + (size_t)m * nrows * Q * 32;
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
      }
      tmem_cta_sync();
      const bool first = m == 0, last = m == nT - 1;
      uint32_t t = t0 + wib;
      int nxL = 0, nxli = 0;
      uint32_t nxoff = 0;
      if (t < t1) { nxL = (int)a.tg.len[t]; nxoff = a.tg.off[t]; nxli = a.tg.lidx[t]; }
      int pre_x = 0;
      bool have_pre = false;
      for (int j = 0; t < t1; t += wpb, j++) {
        const int L = nxL, li = nxli;
        const uint8_t *p = a.tg.dsq + nxoff;
        if (t + wpb < t1) { nxL = (int)a.tg.len[t + wpb]; nxoff = a.tg.off[t + wpb]; nxli = a.tg.lidx[t + wpb]; }
        else nxL = 0;
        if (L == 0 || ((flags >> j) & 1u)) { have_pre = false; continue; }
        const int tjb = a.tjb_li[li];
        const int T = a.thrJ[(size_t)h * a.nL + li];
        const int tjbm = (tjb + tbm) & 0xff;
        const int xB0 = max(base - tjbm, 0);
        const int C = min(255 - bias, min(base + 1 + tec, T + tec));
        const int Cp = a.all_pairs ? 0 : C - xB0;
        if (Cp <= 0) { flags |= 1u << j; have_pre = false; continue; }
        uint32_t w[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) w[q] = 0;
        const uint32_t sc = (uint32_t)a.scJ[(size_t)h * a.nL + li] * 0x10001u;
        uint16_t *brow = bnd_cta + (size_t)(t - t0) * wa.Lpad;
        uint32_t bin_next = (!first && lane < L) ? (uint32_t)__ldcg(brow + lane) : 0u;
        uint32_t lost = 0;  // lane 0: running maximum of what left through the last cell
        if (!have_pre) pre_x = lane < L ? (int)p[lane] : nrows;
        for (int i0 = 0; i0 < L; i0 += 32) {
          const int nr = min(32, L - i0);
          const int myx = pre_x;
          if (i0 + 32 < L) pre_x = i0 + 32 + lane < L ? (int)p[i0 + 32 + lane] : nrows;
          else { pre_x = lane < nxL ? (int)a.tg.dsq[nxoff + lane] : nrows; have_pre = true; }
          // lane r: the value that enters cell 0 of this tile at row i0 + r (high half, where the carry PRMT takes it)
          const uint32_t bin = bin_next << 16;
          if (!first && i0 + 32 < L) bin_next = i0 + 32 + lane < L ? (uint32_t)__ldcg(brow + i0 + 32 + lane) : 0u;
          uint16_t *bout = brow + i0;
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
              uint32_t ds[QS > 0 ? QS : 1];
              if constexpr (QS > 0) msv_load_row<QS>(s_tab + __shfl_sync(PA_FULL, myx, r + g) * SROW, lane, ds);
              // word Q-1 of the left neighbour after the previous row; lane 0 sees the tile's last cell (lane 31, high half):
              // it leaves the tile here -- parked for the next tile (row i0 + r + g of its sweep) or, in the last tile, watched for +inf
              uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
              const uint32_t binr = __shfl_sync(PA_FULL, bin, r + g);
              if (!last && lane == 0) __stcg(bout + r + g, (uint16_t)(tt >> 16));
              lost = __vmaxs2(lost, tt);
              if (lane == 0) tt = binr;
              const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
              uint32_t(&dt)[LW] = buf[(r / G) & 1][g];
              tmem_wait_ld<LW>(dt);
#pragma unroll
              for (int q = Q - 1; q >= 1; q--) w[q] = h2_fma_relu(q < QT ? dt[q < QT ? q : 0] : ds[q >= QT ? q - QT : 0], sc, w[q - 1]);
              w[0] = h2_fma_relu(dt[0], sc, carry);
            }
          }
        }
        const uint32_t mx = row_max_bits<Q>(w);
        bool hot = (mx & 0xffffu) >= 0x7c00u || (mx >> 16) >= 0x7c00u;
        if (last && lane == 0 && (lost >> 16) >= 0x7c00u) hot = true;
        if (__any_sync(PA_FULL, hot)) flags |= 1u << j;
      }
    }
    {
      int j = 0;
      for (uint32_t t = t0 + wib; t < t1; t += wpb, j++)
        if ((flags >> j) & 1u) {
          ncand++;
          if (lane == 0) {
            const unsigned long long idx = atomicAdd(wa.cand_count, 1ull);
            if (idx < wa.cand_cap) {
              Pass1 r;
              r.t = t; r.h = (uint16_t)h; r.xJ = 0;
              wa.cand[idx] = r;
            }
          }
        }
    }
  }
  if (lane == 0 && ncand) atomicAdd(a.n_ssv_cand, ncand);
  tmem_dealloc(tbase, (uint32_t)ncols);
}

// ---------------------------------------------------------------------------------------------
// MSV, exact recurrence for the candidates of wide profiles (and score_all on a wide profile)
// ---------------------------------------------------------------------------------------------
struct MsvExactWideArgs {
  const DevHmm *hmms;
  const uint32_t *tab;       // s16x2 pool; wide profiles: [x][q][lane], q < hm.Q, cell (2*lane + hf)*Q + q
  Targets tg;
  const uint8_t *tjb_li;
  const int16_t *thrJ;
  int nL;
  const Pass1 *cand;         // nullptr: every target against profile only_h
  const unsigned long long *n_cand;
  unsigned long long cand_cap;
  int only_h, all_pairs;
  Pass1 *out;
  unsigned long long *out_count;
  unsigned long long out_cap;
  int maxQ;                  // shared-memory state per warp: maxQ * 32 words
};

static __global__ void __launch_bounds__(128) msv_exact_wide_kernel(MsvExactWideArgs a) {
  extern __shared__ __align__(16) uint32_t s_state[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  uint32_t *w = s_state + (size_t)wib * a.maxQ * 32 + lane;
  const int lanem1 = (lane + 31) & 31;
  const unsigned long long n = a.cand ? min(*a.n_cand, a.cand_cap) : (unsigned long long)a.tg.n;
  for (unsigned long long i = (unsigned long long)blockIdx.x * wpb + wib; i < n; i += (unsigned long long)gridDim.x * wpb) {
    uint32_t t;
    int h;
    if (a.cand) { const Pass1 c = a.cand[i]; t = c.t; h = c.h; }
    else { t = (uint32_t)i; h = a.only_h; }
    const DevHmm &hm = a.hmms[h];
    const int L = (int)a.tg.len[t];
    if (L == 0) continue;
    const int Q = hm.Q, roww = Q * 32;
    const uint8_t *p = a.tg.dsq + a.tg.off[t];
    const int li = a.tg.lidx[t];
    const int tjb = a.tjb_li[li];
    const int T = a.thrJ[(size_t)h * a.nL + li];
    const int bias = hm.bias, base = hm.base, tec = hm.tec;
    const int tjbm = (tjb + hm.tbm) & 0xff;
    const uint32_t *tab = a.tab + hm.msv_off + lane;
    for (int q = 0; q < Q; q++) w[q * 32] = 0;
    int xJ = 0, xB = max(base - tjbm, 0);
    bool overflow = false;
    for (int r = 0; r < L; r++) {
      const uint32_t *row = tab + (size_t)p[r] * roww;
      const uint32_t xBv = (uint32_t)xB * 0x10001u;
      // Q = M/64 + 1: the last striped cell (lane 31, high half, word Q-1) is padding and stays 0, so the carry wraps
      const uint32_t wl = w[(Q - 1) * 32];
      uint32_t left = __byte_perm(__shfl_sync(PA_FULL, wl, lanem1), wl, 0x5432);
      uint32_t rm = 0;
#pragma unroll 4
      for (int q = 0; q < Q; q++) {
        const uint32_t old = w[q * 32];
        const uint32_t nw = __viaddmax_s16x2(__vmaxs2(left, xBv), __ldg(row + q * 32), 0u);
        w[q * 32] = nw;
        rm = __vmaxs2(rm, nw);
        left = old;
      }
      int xE = warp_max16x2(rm);
      if (xE >= 255 - bias) { overflow = true; break; }
      xE = max(xE - tec, 0);
      xJ = max(xJ, xE);
      xB = max(max(base, xJ) - tjbm, 0);
    }
    if (overflow) xJ = -1;
    if ((a.all_pairs || overflow || xJ >= T) && lane == 0) {
      const unsigned long long idx = atomicAdd(a.out_count, 1ull);
      if (idx < a.out_cap) {
        Pass1 r;
        r.t = t; r.h = (uint16_t)h; r.xJ = (int16_t)xJ;
        a.out[idx] = r;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Viterbi filter, one pair per CTA of NW warps.  Thread tid owns the contiguous cells
// k = tid*B + 1 .. (tid+1)*B (packed pass: half-lane 2*tid + hf owns BH cells); tables are [j][tid].
// ---------------------------------------------------------------------------------------------
#define PA_WIDE_MAXW 8   // warps per pair

template <int BH>
__device__ __forceinline__ bool vit16_wide_pair(const DevHmm &hm, const uint32_t *tab, const uint8_t *p, int L, int xw_move, int Tint,
                                                int (*s_red)[PA_WIDE_MAXW], uint32_t (*s_x)[PA_WIDE_MAXW][2], int (*s_tot)[PA_WIDE_MAXW][2]) {
  const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5, NW = NT >> 5;
  const int W = BH * NT;
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
  const uint32_t pslast = __ldg(ps + (BH - 1) * NT + tid);
  const int lanem1 = (lane + 31) & 31;
  uint32_t nbM = FLv, nbI = FLv, nbD = FLv;  // lane 0 of warps > 0: last word of thread tid-1 after the previous row
  for (int i = 0; i < L; i++) {
    const int par = i & 1;
    const uint32_t *emx = em + (size_t)p[i] * W;
    const uint32_t xBv = rep16(xB);
    uint32_t tM = __shfl_sync(PA_FULL, Mx[BH - 1], lanem1), tI = __shfl_sync(PA_FULL, Ix[BH - 1], lanem1),
             tD = __shfl_sync(PA_FULL, Dx[BH - 1], lanem1);
    if (lane == 0) { tM = nbM; tI = nbI; tD = nbD; }
    uint32_t mL = __byte_perm(tM, Mx[BH - 1], 0x5432), iL = __byte_perm(tI, Ix[BH - 1], 0x5432), dL = __byte_perm(tD, Dx[BH - 1], 0x5432);
    if (tid == 0) {  // nothing left of cell 1
      mL = (mL & 0xffff0000u) | (FLv & 0xffffu);
      iL = (iL & 0xffff0000u) | (FLv & 0xffffu);
      dL = (dL & 0xffff0000u) | (FLv & 0xffffu);
    }
    uint32_t rowmax = FLv;
#pragma unroll
    for (int j = BH - 1; j >= 0; j--) {
      const int c = j * NT + tid;
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
    const int wmax = __reduce_max_sync(PA_FULL, max(lo16(rowmax), hi16(rowmax)));
    if (lane == 0) s_red[par][wv] = wmax;
    if (lane == 31) { s_x[par][wv][0] = Mx[BH - 1]; s_x[par][wv][1] = Ix[BH - 1]; }
    __syncthreads();
    int xE = s_red[par][0];
    for (int v = 1; v < NW; v++) xE = max(xE, s_red[par][v]);
    if (xE + xw_E >= Tint) return true;  // the same value in every thread: the whole CTA leaves together
    uint32_t tMc = __shfl_sync(PA_FULL, Mx[BH - 1], lanem1);
    if (lane == 0 && wv) { nbM = s_x[par][wv - 1][0]; nbI = s_x[par][wv - 1][1]; }
    if (lane == 0) tMc = nbM;
    uint32_t mcL = __byte_perm(tMc, Mx[BH - 1], 0x5432);
    if (tid == 0) mcL = (mcL & 0xffff0000u) | (FLv & 0xffffu);
    // ---- D: serial inside each half-lane, max-plus scan over the half-lanes of the warp, then over the warps
    uint32_t aloc = FLv;
#pragma unroll
    for (int j = 0; j < BH; j++) {
      const uint2 Dq = __ldg(trD + j * NT + tid);
      const uint32_t mleft = j ? Mx[j > 0 ? j - 1 : 0] : mcL;
      const uint32_t t = __viaddmax_s16x2(mleft, Dq.x, FLv);
      aloc = __viaddmax_s16x2(aloc, Dq.y, t);
      Dx[j] = aloc;
    }
    const int Alo = lo16(aloc), Ahi0 = hi16(aloc), Slo = lo16(pslast), Shi0 = hi16(pslast);
    int LA = max(Ahi0, max(Alo + Shi0, PA_V16_FLOOR)), LS = max(Slo + Shi0, PA_V16_FLOOR);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a2 = __shfl_up_sync(PA_FULL, LA, o), s2 = __shfl_up_sync(PA_FULL, LS, o);
      if (lane >= o) {
        LA = max(LA, max(a2 + LS, PA_V16_FLOOR));
        LS = max(LS + s2, PA_V16_FLOOR);
      }
    }
    if (lane == 31) { s_tot[par][wv][0] = LA; s_tot[par][wv][1] = LS; }
    __syncthreads();
    int cA = PA_V16_FLOOR, cS = 0;  // D leaving the last half-lane of warp wv-1, given everything to its left
    for (int v = 0; v < wv; v++) {
      const int A = s_tot[par][v][0], S = s_tot[par][v][1];
      cA = max(A, max(cA + S, PA_V16_FLOOR));
      cS = max(cS + S, PA_V16_FLOOR);
    }
    if (wv) LA = max(LA, max(cA + LS, PA_V16_FLOOR));
    int din_lo = __shfl_up_sync(PA_FULL, LA, 1);
    if (lane == 0) din_lo = cA;
    const int din_hi = max(Alo, max(din_lo + Slo, PA_V16_FLOOR));
    const uint32_t din = (uint32_t)(din_lo & 0xffff) | ((uint32_t)(din_hi & 0xffff) << 16);
#pragma unroll
    for (int j = 0; j < BH; j++) Dx[j] = __viaddmax_s16x2(din, __ldg(ps + j * NT + tid), Dx[j]);
    if (lane == 0) nbD = (uint32_t)(cA & 0xffff) << 16;  // an upper bound of the D that leaves thread tid-1 (floors only raise)
    xJ = max(xJ, max(xE + xw_E, PA_V16_FLOOR));
    xB = max(max(xJ + xw_move, xN + xw_move), PA_V16_FLOOR);
  }
  return false;
}

template <int BH>
static __global__ void __launch_bounds__(128, 2) vit16_wide_kernel(Vit16Args a) {
  __shared__ int s_red[2][PA_WIDE_MAXW];
  __shared__ uint32_t s_x[2][PA_WIDE_MAXW][2];
  __shared__ int s_tot[2][PA_WIDE_MAXW][2];
  unsigned long long nneed = 0;
  for (unsigned long long i = a.begin + blockIdx.x; i < a.end; i += gridDim.x) {
    const Pass2 r = a.in[i];
    const DevHmm &hm = a.hmms[r.h];
    bool need = true;
    if (r.flags & 1) {
      const int L = (int)a.tg.len[r.t];
      const int xw_move = a.lt.xw_move[a.tg.lidx[r.t]];
      const float need_sc = (float)((double)hm.thrF2v * PA_LN2) + r.filtersc + 3.0f;
      const float tf = need_sc * hm.scale_w + (float)hm.base_w - (float)xw_move;
      if (tf > (float)(PA_V16_FLOOR + 8) && tf < 32000.0f) {
        const int Tint = (int)floorf(tf) - 2;
        if (Tint + hm.v16_guard < 32767) need = vit16_wide_pair<BH>(hm, a.tab, a.tg.dsq + a.tg.off[r.t], L, xw_move, Tint, s_red, s_x, s_tot);
      }
    }
    if (threadIdx.x == 0) { a.need[i] = need ? 1 : 0; nneed += need ? 1 : 0; }
    __syncthreads();
  }
  if (threadIdx.x == 0 && nneed) atomicAdd(a.n_need, nneed);
}

template <int BT>
__device__ __forceinline__ int vit_wide_pair(const DevHmm &hm, const int *tab, const uint8_t *p, int L, int xw_move,
                                             int (*s_red)[PA_WIDE_MAXW], int (*s_x)[PA_WIDE_MAXW][2], int (*s_tot)[PA_WIDE_MAXW][2]) {
  const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5, NW = NT >> 5;
  const int W = BT * NT;
  const int4 *trA = reinterpret_cast<const int4 *>(tab + hm.vit_off);
  const int2 *trI = reinterpret_cast<const int2 *>(trA + W);
  const int2 *trD = trI + W;
  const int *ps = reinterpret_cast<const int *>(trD + W);
  const int *em = ps + W;
  int Mx[BT], Ix[BT], Dx[BT];
#pragma unroll
  for (int j = 0; j < BT; j++) { Mx[j] = PA_NEG16; Ix[j] = PA_NEG16; Dx[j] = PA_NEG16; }
  const int xN = hm.base_w, xw_E = hm.xw_E;
  int xB = max(xN + xw_move, PA_NEG16), xJ = PA_NEG16, xC = PA_NEG16;
  const int pslast = __ldg(ps + (BT - 1) * NT + tid);
  int nbM = PA_NEG16, nbI = PA_NEG16, nbD = PA_NEG16;  // lane 0 of warps > 0: last cell of thread tid-1 after the previous row
  for (int i = 0; i < L; i++) {
    const int par = i & 1;
    const int *emx = em + (size_t)p[i] * W;
    int mL = __shfl_up_sync(PA_FULL, Mx[BT - 1], 1), iL = __shfl_up_sync(PA_FULL, Ix[BT - 1], 1), dL = __shfl_up_sync(PA_FULL, Dx[BT - 1], 1);
    if (lane == 0) { mL = nbM; iL = nbI; dL = nbD; }
    int rowmax = PA_NEG16;
#pragma unroll
    for (int j = BT - 1; j >= 0; j--) {
      const int c = j * NT + tid;
      const int4 A = __ldg(trA + c);
      const int2 Iq = __ldg(trI + c);
      const int e = __ldg(emx + c);
      const int pm = j ? Mx[j > 0 ? j - 1 : 0] : mL, pi = j ? Ix[j > 0 ? j - 1 : 0] : iL, pd = j ? Dx[j > 0 ? j - 1 : 0] : dL;
      int sv = __viaddmax_s32(xB, A.x, PA_NEG16);
      sv = __viaddmax_s32(pm, A.y, sv);
      sv = __viaddmax_s32(pi, A.z, sv);
      sv = __viaddmax_s32(pd, A.w, sv);
      sv = __viaddmax_s32(sv, e, PA_NEG16);
      int iv = __viaddmax_s32(Mx[j], Iq.x, PA_NEG16);
      iv = __viaddmax_s32(Ix[j], Iq.y, iv);
      Mx[j] = sv;
      Ix[j] = iv;
      rowmax = max(rowmax, sv);
      if ((j & 3) == 0) asm volatile("" ::: "memory");
    }
    const int wmax = __reduce_max_sync(PA_FULL, rowmax);
    if (lane == 0) s_red[par][wv] = wmax;
    if (lane == 31) { s_x[par][wv][0] = Mx[BT - 1]; s_x[par][wv][1] = Ix[BT - 1]; }
    __syncthreads();
    int xE = s_red[par][0];
    for (int v = 1; v < NW; v++) xE = max(xE, s_red[par][v]);
    if (xE >= 32767) return 32767;
    int mcL = __shfl_up_sync(PA_FULL, Mx[BT - 1], 1);
    if (lane == 0) {
      if (wv) { nbM = s_x[par][wv - 1][0]; nbI = s_x[par][wv - 1][1]; mcL = nbM; }
      else mcL = PA_NEGBIG;
    }
    int aloc = PA_NEGBIG;
#pragma unroll
    for (int j = 0; j < BT; j++) {
      const int2 Dq = __ldg(trD + j * NT + tid);
      const int mleft = j ? Mx[j > 0 ? j - 1 : 0] : mcL;
      aloc = max(max(mleft + Dq.x, aloc + Dq.y), PA_NEGBIG);
      Dx[j] = aloc;
    }
    int LA = aloc, LS = pslast;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a2 = __shfl_up_sync(PA_FULL, LA, o), s2 = __shfl_up_sync(PA_FULL, LS, o);
      if (lane >= o) {
        LA = max(LA, max(a2 + LS, PA_NEGBIG));
        LS = max(LS + s2, PA_NEGBIG);
      }
    }
    if (lane == 31) { s_tot[par][wv][0] = LA; s_tot[par][wv][1] = LS; }
    __syncthreads();
    int cA = PA_NEGBIG, cS = 0;
    for (int v = 0; v < wv; v++) {
      const int A = s_tot[par][v][0], S = s_tot[par][v][1];
      cA = max(A, max(cA + S, PA_NEGBIG));
      cS = max(cS + S, PA_NEGBIG);
    }
    if (wv) LA = max(LA, max(cA + LS, PA_NEGBIG));
    int din = __shfl_up_sync(PA_FULL, LA, 1);
    if (lane == 0) din = cA;
#pragma unroll
    for (int j = 0; j < BT; j++) Dx[j] = __vimax3_s32(Dx[j], din + __ldg(ps + j * NT + tid), PA_NEG16);
    if (lane == 0 && wv) nbD = max(cA, PA_NEG16);  // = D of the last cell of thread tid-1 (the floor is applied once, at the end)
    xC = max(xC, max(xE + xw_E, PA_NEG16));
    xJ = max(xJ, max(xE + xw_E, PA_NEG16));
    xB = max(max(xJ + xw_move, xN + xw_move), PA_NEG16);
  }
  return xC;
}

template <int BT>
static __global__ void __launch_bounds__(256, 1) vit_wide_kernel(VitArgs a) {
  __shared__ int s_red[2][PA_WIDE_MAXW];
  __shared__ int s_x[2][PA_WIDE_MAXW][2];
  __shared__ int s_tot[2][PA_WIDE_MAXW][2];
  for (unsigned long long i = a.begin + blockIdx.x; i < a.end; i += gridDim.x) {
    if (a.need && !a.need[i]) continue;
    const Pass2 r = a.in[i];
    const DevHmm &hm = a.hmms[r.h];
    if (!(r.flags & 1) && !a.all_pairs) {
      if (threadIdx.x == 0) {
        unsigned long long idx = atomicAdd(a.out_count, 1ull);
        if (idx < a.out_cap) a.out[idx] = r;
      }
      continue;
    }
    const int L = (int)a.tg.len[r.t];
    const int xw_move = a.lt.xw_move[a.tg.lidx[r.t]];
    const int xC = vit_wide_pair<BT>(hm, a.tab, a.tg.dsq + a.tg.off[r.t], L, xw_move, s_red, s_x, s_tot);
    __syncthreads();
    if (a.raw_out && threadIdx.x == 0) a.raw_out[i] = xC;
    if (a.all_pairs) continue;
    const float vfsc = vit_score_from_xC(hm, xC, xw_move);
    const float seq = (float)((double)(vfsc - r.filtersc) / PA_LN2);
    if (seq >= hm.thrF2v && threadIdx.x == 0) {
      unsigned long long idx = atomicAdd(a.out_count, 1ull);
      if (idx < a.out_cap) a.out[idx] = r;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward filter, run-time block size (B = ceil(M/32) > 64)
//
// The arithmetic and its order are fwd_pair_t's (kernels.cuh), so scores are bit-identical to the oracle's; what
// is specific to wide profiles is where the data sits.  One warp per CTA; the M, I, D rows of the pair live in
// shared memory as [j][lane] blocks updated in place (3 x B x 128 bytes: as many CTAs per SM as that allows), the
// D->D prefix products P(j) -- a property of the profile, not of the row -- are computed once per profile into a
// per-CTA global row (each lane reads back only what it wrote), and the table reads of a row are issued
// PA_FWDW_U cells deep before the dependent arithmetic, because with 2-6 resident warps per SM the L2 latency of
// the transition/emission tables is what the kernel waits on (profiles/r2_ncu_config3_vit_fwd_domain.md: the
// previous global-scratch version moved 232 GB of DRAM traffic per launch for its seven rows per warp).
// ---------------------------------------------------------------------------------------------
#define PA_FWDW_U 8
struct FwdWideArgs {
  FwdTArgs ft;
  float *scratch;     // per CTA: per_cta floats (P row of the profile in hand)
  size_t per_cta;
};

static __global__ void __launch_bounds__(32) fwd_wide_kernel(FwdWideArgs w) {
  extern __shared__ __align__(16) float s_fww[];
  constexpr int U = PA_FWDW_U;
  const FwdTArgs &a = w.ft;
  const int lane = threadIdx.x;
  float *Pg = w.scratch + (size_t)blockIdx.x * w.per_cta + lane;  // P(j) = Pg[j * 32]
  int last_h = -1;
  float LPtot = 1.0f;
  for (unsigned long long i = a.begin + blockIdx.x; i < a.end; i += gridDim.x) {
    const Pass2 r = a.f.in[i];
    const DevHmm &hm = a.f.hmms[r.h];
    const FwdTabs T = fwd_tabs(hm, a.f.tab);
    const int B = T.B, W = T.W;
    const int L = (int)a.f.tg.len[r.t];
    const uint8_t *p = a.f.tg.dsq + a.f.tg.off[r.t];
    const int nv = min(B, max(0, T.M - lane * B));  // valid cells of this lane
    float *Mr = s_fww + lane, *Ir = Mr + W, *Dr = Ir + W;
    if ((int)r.h != last_h) {  // running products of DD along the lane's block, in fwd_pair_t's order
      float LP = 1.0f;
      for (int j = 0; j < nv; j++) {
        const float y = __ldg(&T.trB[j * 32 + lane].y);
        LP = j ? __fmul_rn(LP, y) : y;
        Pg[j * 32] = LP;
      }
      LPtot = LP;
      last_h = (int)r.h;
    }
    for (int j = 0; j < B; j++) { Mr[j * 32] = 0.0f; Ir[j * 32] = 0.0f; Dr[j * 32] = 0.0f; }
    const XF xf = xf_config(L, 1);
    float xN = 1.0f, xB = xf.tNB, xE = 0.0f, xJ = 0.0f, xC = 0.0f;
    int etot = 0;
    int xnext = L > 0 ? (int)p[0] : 0;
    for (int row = 0; row < L; row++) {
      const float *em = T.em + (size_t)xnext * W + lane;
      if (row + 1 < L) xnext = (int)p[row + 1];
      float mL = __shfl_up_sync(PA_FULL, Mr[(B - 1) * 32], 1), iL = __shfl_up_sync(PA_FULL, Ir[(B - 1) * 32], 1),
            dL = __shfl_up_sync(PA_FULL, Dr[(B - 1) * 32], 1);
      if (lane == 0) mL = iL = dL = 0.0f;
      // M and I, in place: descending j reads the previous row's j-1 before it is overwritten
      for (int jb = B - 1; jb >= 0; jb -= U) {
        float4 A[U];
        float2 Bz[U];
        float e[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb - u;
          if (j >= 0 && j < nv) {
            const int c = j * 32 + lane;
            A[u] = __ldg(T.trA + c);
            const float4 q = __ldg(T.trB + c);
            Bz[u] = make_float2(q.z, q.w);
            e[u] = __ldg(em + j * 32);
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb - u;
          if (j >= 0 && j < nv) {
            const int jm = j > 0 ? j - 1 : 0;
            const float pm = j ? Mr[jm * 32] : mL, pi = j ? Ir[jm * 32] : iL, pd = j ? Dr[jm * 32] : dL;
            float t = __fmul_rn(xB, A[u].x);
            t = __fadd_rn(t, __fmul_rn(pm, A[u].y));
            t = __fadd_rn(t, __fmul_rn(pi, A[u].z));
            t = __fadd_rn(t, __fmul_rn(pd, A[u].w));
            const float ni = __fadd_rn(__fmul_rn(Mr[j * 32], Bz[u].x), __fmul_rn(Ir[j * 32], Bz[u].y));
            Mr[j * 32] = __fmul_rn(t, e[u]);
            Ir[j * 32] = ni;
          }
        }
      }
      // D chain: in-lane serial prefix, scan across lanes, fix-up
      float mcL = __shfl_up_sync(PA_FULL, Mr[(B - 1) * 32], 1);
      if (lane == 0) mcL = 0.0f;
      float LA = 0.0f;
      for (int jb = 0; jb < B; jb += U) {
        float2 Bx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb + u;
          if (j < nv) {
            const float4 q = __ldg(T.trB + j * 32 + lane);
            Bx[u] = make_float2(q.x, q.y);
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb + u;
          if (j < nv) {
            const float av = __fmul_rn(j ? Mr[(j > 0 ? j - 1 : 0) * 32] : mcL, Bx[u].x);
            LA = j ? __fadd_rn(av, __fmul_rn(LA, Bx[u].y)) : av;
            Dr[j * 32] = LA;
          }
        }
      }
      float LP = nv > 0 ? LPtot : 1.0f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float a2 = __shfl_up_sync(PA_FULL, LA, o), p2 = __shfl_up_sync(PA_FULL, LP, o);
        if (lane >= o) { LA = __fadd_rn(LA, __fmul_rn(LP, a2)); LP = __fmul_rn(LP, p2); }
      }
      float cin = __shfl_up_sync(PA_FULL, LA, 1);
      if (lane == 0) cin = 0.0f;
      float sacc = 0.0f;
      for (int jb = 0; jb < B; jb += U) {
        float P[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb + u;
          if (j < nv) P[u] = Pg[j * 32];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = jb + u;
          if (j < nv) {
            const float d = __fadd_rn(Dr[j * 32], __fmul_rn(P[u], cin));
            Dr[j * 32] = d;
            sacc = __fadd_rn(sacc, __fadd_rn(Mr[j * 32], d));
          }
        }
      }
      xE = lane_sum(sacc);
      xN = __fmul_rn(xN, xf.tNN);
      xJ = __fadd_rn(__fmul_rn(xJ, xf.tJJ), __fmul_rn(xE, xf.tEJ));
      xC = __fadd_rn(__fmul_rn(xC, xf.tCC), __fmul_rn(xE, xf.tEC));
      xB = __fadd_rn(__fmul_rn(xJ, xf.tJB), __fmul_rn(xN, xf.tNB));
      if (xE > PA_RESCALE) {
        const int e = ilogbf(xE);
        for (int j = 0; j < B; j++) {
          Mr[j * 32] = scalbnf(Mr[j * 32], -e); Ir[j * 32] = scalbnf(Ir[j * 32], -e); Dr[j * 32] = scalbnf(Dr[j * 32], -e);
        }
        xN = scalbnf(xN, -e); xJ = scalbnf(xJ, -e); xC = scalbnf(xC, -e); xB = scalbnf(xB, -e); xE = scalbnf(xE, -e);
        etot += e;
      }
    }
    const float mant = __fmul_rn(xC, xf.tCT);
    const float fwdsc = (float)((double)etot * PA_LN2 + log((double)mant));
    const float seq = (float)((double)(fwdsc - r.filtersc) / PA_LN2);
    if (lane == 0 && (a.f.all_pairs || seq >= hm.thrF3)) {
      Pass4 o;
      o.t = r.t; o.h = r.h; o.pad = 0; o.filtersc = r.filtersc; o.fexp = etot; o.fmant = mant;
      if (a.f.all_pairs) a.f.out[i] = o;
      else {
        unsigned long long idx = atomicAdd(a.f.out_count, 1ull);
        if (idx < a.f.out_cap) a.f.out[idx] = o;
      }
    }
  }
}

}  // namespace pa
