// domain.cuh -- survivor stage on the GPU: Backward parser, posterior domain definition,
// envelope rescoring (unihit Forward/Backward), posterior decoding, optimal-accuracy alignment
// with traceback, and null2 by expectation.
//
// One warp per Forward-filter survivor, persistent warps pulling work from an atomic counter.
// Each resident warp owns a global-memory scratch slab (DP rows, per-row specials, the two full
// (Ld+1) x W matrices of an envelope).  Lane l owns model cells k = l*B+1..(l+1)*B and only ever
// touches its own cells in memory; cross-lane traffic is shuffles only.  Serial pieces (region
// scan, traceback) run on lane 0.  Float evaluation order is the canonical one shared with the
// oracle (DESIGN.md "Canonical float order"), so regions, envelopes and tracebacks are bit-exact.
//
// Regions that trip the multi-domain test (rt3) are resolved as HMMER does: 200 stochastic tracebacks of
// the region's multihit Forward matrix (one trace per lane, own xorshift64* stream per trace index),
// single-linkage clustering of the sampled segments (label propagation over all lanes), envelopes from
// the clusters seen in >= 25 % of the traces.  Same arithmetic and order as oracle/p7_domain.c.
#pragma once
#include "kernels.cuh"

namespace pa {

#ifndef PA_DOM_MINB
#define PA_DOM_MINB 4   // resident 128-thread blocks per SM the register budget must allow (4: 128 registers, 16 warps)
#endif
#define PA_MAXREG 64
#define PA_NSAMPLES 200   // stochastic tracebacks per multi-domain region (HMMER: 200)
#define PA_MAXSEG 16     // domain segments per sampled trace; more is an error (status bit 1), never a silent drop
#define PA_MAXENV 16     // envelopes per multi-domain region; same

struct DomRec {
  uint32_t t;
  uint16_t h;
  uint16_t multi;
  int surv, dom_idx;
  int ienv, jenv, iali, jali, hmmfrom, hmmto;
  int env_exp;
  float env_mant;
  float oasc;
  int ali_off, ali_len, res_off;
  float null2[32];
};

struct DomainBufs {
  float *scratch = nullptr;
  size_t scratch_floats = 0;
  DomRec *recs = nullptr;
  size_t recs_cap = 0;
  char *model = nullptr, *aligned = nullptr;
  uint8_t *res = nullptr;
  size_t pool_cap = 0, res_cap = 0;
  char *cons = nullptr;  // consensus letters of all profiles
  size_t cons_cap = 0;
  void release() {
    if (scratch) cudaFree(scratch);
    if (recs) cudaFree(recs);
    if (model) cudaFree(model);
    if (aligned) cudaFree(aligned);
    if (res) cudaFree(res);
    if (cons) cudaFree(cons);
    scratch = nullptr; recs = nullptr; model = aligned = nullptr; res = nullptr; cons = nullptr;
    scratch_floats = recs_cap = pool_cap = res_cap = cons_cap = 0;
  }
};

struct SlabLayout {
  size_t sf, sb, ef, eb, lin, reg, trk, tri, trs, seg, mat, total;  // offsets in floats
  size_t msz;
};
__host__ __device__ inline SlabLayout slab_layout(int maxL, int maxW, int maxM) {
  SlabLayout s;
  const size_t L1 = (size_t)maxL + 1, tr = (size_t)maxL + maxM + 16;
  size_t o = 8 * (size_t)maxW;
  s.sf = o; o += 5 * L1;
  s.sb = o; o += 5 * L1;
  s.ef = o; o += L1;
  s.eb = o; o += L1;
  s.lin = o; o += 11 * L1;       // btot etot mocc ppN ppJ ppC oE oN oJ oB oC
  s.reg = o; o += 3 * PA_MAXREG;
  s.trk = o; o += tr;
  s.tri = o; o += tr;
  s.trs = o; o += (tr + 3) / 4;
  s.seg = o; o += 10 * (size_t)PA_NSAMPLES * PA_MAXSEG + PA_NSAMPLES + 4 * PA_MAXENV + 64;  // multi-domain regions: segraw 4x | segc 5x | lab 1x | nper | envl
  o = (o + 3) & ~(size_t)3;
  s.mat = o;
  s.msz = L1 * (size_t)maxW;
  s.total = o + 6 * s.msz;
  return s;
}

// HMMER's p7_alidisplay_EncodePostProb: '0'..'9', '*' from 0.95
__device__ __forceinline__ char pp_char(float p) {
  const double q = (double)p + 0.05;
  return q >= 1.0 ? '*' : (char)((int)(q * 10.0) + '0');
}

struct DomArgs {
  const DevHmm *hmms;
  const float *tab;
  Targets tg;
  const Pass4 *in;
  unsigned long long n_in;
  unsigned long long *work_counter;   // next survivor
  unsigned long long *rec_count, *pool_count, *res_count;
  unsigned long long rec_cap, pool_cap, res_cap;
  DomRec *recs;
  char *model, *aligned;
  uint8_t *res;
  const char *cons;                   // per profile: cons_off[h] .. +M+1 (1-based)
  const unsigned long long *cons_off;
  float *scratch;
  size_t slab_floats;                 // per warp
  int maxL, maxW, maxM;
  int do_null2;
  int abc;
  int *status;                        // bit0: capacity overflow
};

// One backward row. Mn/In: row i+1 (zeros for the last row); em: emission row of residue x(i+1).
// last: row L (no next residue; sp preset by the caller). Arrays [j*32+lane].
__device__ void bwd_row(const FwdTabs &T, const float *em, const float *Mn, const float *In, const XF &xf,
                        const Specials &sn, Specials &sp, float *Mc, float *Ic, float *Dc, float *we, float *wP,
                        int lane, bool last) {
  const int B = T.B;
  const int nv = min(B, max(0, T.M - lane * B));
  if (!last) {
    float s = 0.0f;
    for (int j = 0; j < nv; j++) {
      const int c = j * 32 + lane;
      const float w = __fmul_rn(Mn[c], em[c]);
      we[c] = w;
      s = __fadd_rn(s, __fmul_rn(w, T.bwB[c].w));
    }
    sp.xB = lane_sum(s);
    sp.xJ = __fadd_rn(__fmul_rn(sp.xB, xf.tJB), __fmul_rn(sn.xJ, xf.tJJ));
    sp.xC = __fmul_rn(sn.xC, xf.tCC);
    sp.xE = __fadd_rn(__fmul_rn(sp.xJ, xf.tEJ), __fmul_rn(sp.xC, xf.tEC));
    sp.xN = __fadd_rn(__fmul_rn(sp.xB, xf.tNB), __fmul_rn(sn.xN, xf.tNN));
  } else {
    for (int j = 0; j < B; j++) we[j * 32 + lane] = 0.0f;
  }
  const float xE = sp.xE;
  float weR = __shfl_down_sync(PA_FULL, we[lane], 1);
  if (lane == 31) weR = 0.0f;
  float LA = 0.0f, LP = 1.0f;
  for (int j = nv - 1; j >= 0; j--) {
    const int c = j * 32 + lane;
    const float4 B4 = T.bwB[c];
    const float wnext = (j + 1 < B) ? we[c + 32] : weR;
    const float cc = __fadd_rn(__fmul_rn(wnext, B4.y), xE);
    if (j == nv - 1) { LA = cc; LP = B4.z; }
    else { LA = __fadd_rn(cc, __fmul_rn(LA, B4.z)); LP = __fmul_rn(LP, B4.z); }
    Dc[c] = LA;
    wP[c] = LP;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float a2 = __shfl_down_sync(PA_FULL, LA, o), p2 = __shfl_down_sync(PA_FULL, LP, o);
    if (lane + o < 32) { LA = __fadd_rn(LA, __fmul_rn(LP, a2)); LP = __fmul_rn(LP, p2); }
  }
  float cin = __shfl_down_sync(PA_FULL, LA, 1);
  if (lane == 31) cin = 0.0f;
  for (int j = 0; j < nv; j++) {
    const int c = j * 32 + lane;
    Dc[c] = __fadd_rn(Dc[c], __fmul_rn(wP[c], cin));
  }
  float dR = __shfl_down_sync(PA_FULL, Dc[lane], 1);
  if (lane == 31) dR = 0.0f;
  for (int j = 0; j < nv; j++) {
    const int c = j * 32 + lane;
    const float4 A4 = T.bwA[c], B4 = T.bwB[c];
    const float wnext = (j + 1 < B) ? we[c + 32] : weR;
    const float dnext = (j + 1 < B) ? Dc[c + 32] : dR;
    float t = __fadd_rn(__fmul_rn(wnext, A4.x), __fmul_rn(In[c], A4.y));
    t = __fadd_rn(t, __fmul_rn(dnext, A4.z));
    Mc[c] = __fadd_rn(t, xE);
    Ic[c] = __fadd_rn(__fmul_rn(wnext, A4.w), __fmul_rn(In[c], B4.x));
  }
}

__device__ __forceinline__ float warp_fmax(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PA_FULL, v, o));
  return v;
}

// Backward over p[0..L-1]. If mx != nullptr rows are stored at mx{M,I,D} + i*W (i = 0..L), else two
// alternating rows inside `rows` (6W) are used. sp/sexp get rows 0..L. rows also provides we,wP,zero.
__device__ void backward_pass(const FwdTabs &T, const uint8_t *p, int L, int Lcfg, int multihit, float *rows,
                              float *mxM, float *mxI, float *mxD, Specials *sp, int *sexp, int lane) {
  const int W = T.W, B = T.B;
  float *we = rows + 6 * W, *wP = rows + 7 * W;
  const XF xf = xf_config(Lcfg, multihit);
  Specials sn, sc;
  sn.xC = xf.tCT; sn.xE = __fmul_rn(sn.xC, xf.tEC); sn.xJ = 0.0f; sn.xB = 0.0f; sn.xN = 0.0f;
  int etot = 0;
  float *Mn, *In;
  // row L
  float *zero = rows + 5 * W;  // a zero row (kept zero: D rows of the parser live elsewhere)
  {
    float *Mc = mxM ? mxM + (size_t)L * W : rows, *Ic = mxI ? mxI + (size_t)L * W : rows + W,
          *Dc = mxD ? mxD + (size_t)L * W : rows + 2 * W;
    for (int j = 0; j < B; j++) { zero[j * 32 + lane] = 0.0f; Mc[j * 32 + lane] = 0.0f; Ic[j * 32 + lane] = 0.0f; Dc[j * 32 + lane] = 0.0f; }
    sc = sn;
    bwd_row(T, nullptr, zero, zero, xf, sn, sc, Mc, Ic, Dc, we, wP, lane, true);
    Mn = Mc; In = Ic;
  }
  if (lane == 0) { sp[L] = sn; sexp[L] = 0; }
  for (int i = L - 1; i >= 0; i--) {
    float *Mc, *Ic, *Dc;
    if (mxM) { Mc = mxM + (size_t)i * W; Ic = mxI + (size_t)i * W; Dc = mxD + (size_t)i * W; }
    else {
      const int par = (L - i) & 1;  // alternate between row sets 0 and 1
      Mc = rows + (par ? 3 * W : 0); Ic = Mc + W; Dc = rows + 2 * W;  // D shared: only used inside a row
    }
    if (mxM) for (int j = 0; j < B; j++) { Mc[j * 32 + lane] = 0.0f; Ic[j * 32 + lane] = 0.0f; Dc[j * 32 + lane] = 0.0f; }
    const float *em = T.em + (size_t)p[i] * W;
    bwd_row(T, em, Mn, In, xf, sn, sc, Mc, Ic, Dc, we, wP, lane, false);
    float rm = 0.0f;
    for (int j = 0; j < B; j++) rm = fmaxf(rm, Mc[j * 32 + lane]);
    rm = warp_fmax(rm);
    if (rm > PA_RESCALE) {
      const int e = ilogbf(rm);
      scale_row(Mc, Ic, Dc, B, lane, e);
      sc.xN = scalbnf(sc.xN, -e); sc.xJ = scalbnf(sc.xJ, -e); sc.xC = scalbnf(sc.xC, -e);
      sc.xB = scalbnf(sc.xB, -e); sc.xE = scalbnf(sc.xE, -e);
      etot += e;
    }
    if (lane == 0) { sp[i] = sc; sexp[i] = etot; }
    sn = sc;
    Mn = Mc; In = Ic;
  }
}

__device__ __forceinline__ int cell_of_k(int k, int B) { return ((k - 1) % B) * 32 + (k - 1) / B; }

// Forward over sub[0..Ld-1] with the full matrix stored (rows 0..Ld at F* + i*W), specials and exponent
// totals per row. xf: multihit or unihit configuration for the FULL target length.
__device__ void forward_full(const FwdTabs &T, const uint8_t *sub, int Ld, const XF &xf, float *rows, float *FM, float *FI,
                             float *FD, Specials *sf, int *ef, int lane, int *Ze_out, float *Zm_out) {
  const int W = T.W, B = T.B;
  for (int c = lane; c < 8 * W; c += 32) rows[c] = 0.0f;
  for (int j = 0; j < B; j++) { FM[j * 32 + lane] = 0.0f; FI[j * 32 + lane] = 0.0f; FD[j * 32 + lane] = 0.0f; }
  int Ze = 0;
  float xN = 1.0f, xB = xf.tNB, xE = 0.0f, xJ = 0.0f, xC = 0.0f;
  float *wP = rows + 6 * W;
  if (lane == 0) { Specials s0; s0.xE = 0; s0.xN = xN; s0.xJ = 0; s0.xB = xB; s0.xC = 0; sf[0] = s0; ef[0] = 0; }
  for (int i = 1; i <= Ld; i++) {
    float *Mp = FM + (size_t)(i - 1) * W, *Ip = FI + (size_t)(i - 1) * W, *Dp = FD + (size_t)(i - 1) * W;
    float *Mc = FM + (size_t)i * W, *Ic = FI + (size_t)i * W, *Dc = FD + (size_t)i * W;
    for (int j = 0; j < B; j++) { Mc[j * 32 + lane] = 0.0f; Ic[j * 32 + lane] = 0.0f; Dc[j * 32 + lane] = 0.0f; }
    xE = fwd_row(T, sub[i - 1], Mp, Ip, Dp, xB, Mc, Ic, Dc, wP, lane);
    xN = __fmul_rn(xN, xf.tNN);
    xJ = __fadd_rn(__fmul_rn(xJ, xf.tJJ), __fmul_rn(xE, xf.tEJ));
    xC = __fadd_rn(__fmul_rn(xC, xf.tCC), __fmul_rn(xE, xf.tEC));
    xB = __fadd_rn(__fmul_rn(xJ, xf.tJB), __fmul_rn(xN, xf.tNB));
    if (xE > PA_RESCALE) {
      const int e = ilogbf(xE);
      scale_row(Mc, Ic, Dc, B, lane, e);
      xN = scalbnf(xN, -e); xJ = scalbnf(xJ, -e); xC = scalbnf(xC, -e); xB = scalbnf(xB, -e); xE = scalbnf(xE, -e);
      Ze += e;
    }
    if (lane == 0) { Specials s; s.xE = xE; s.xN = xN; s.xJ = xJ; s.xB = xB; s.xC = xC; sf[i] = s; ef[i] = Ze; }
  }
  *Ze_out = Ze;
  *Zm_out = __fmul_rn(xC, xf.tCT);
}

// ---- multi-domain regions (oracle/p7_domain.c: p7d_sample_trace, p7d_linked, p7d_resolve_multidomain) ----
__device__ __forceinline__ float md_uniform(unsigned long long &x) {
  x ^= x >> 12;
  x ^= x << 25;
  x ^= x >> 27;
  return __fmul_rn((float)((x * 0x2545F4914F6CDD1Dull) >> 40), 1.0f / 16777216.0f);
}
__device__ __forceinline__ int md_choose(const float *pr, int n, float u) {
  float tot = 0.0f;
  for (int i = 0; i < n; i++) tot = __fadd_rn(tot, pr[i]);
  const float r = __fmul_rn(u, tot);
  float cum = 0.0f;
  int last = 0;
  for (int i = 0; i < n; i++) {
    if (pr[i] > 0.0f) last = i;
    cum = __fadd_rn(cum, pr[i]);
    if (r < cum) return i;
  }
  return last;
}

// One stochastic traceback, run by ONE lane. Writes up to PA_MAXSEG segments {i, j, k, m}; returns their number.
__device__ int md_sample_trace(const FwdTabs &T, const float *FM, const float *FI, const float *FD, const Specials *sf,
                               const int *ef, int Lr, const XF &xf, int t, int *seg4, int *status) {
  const int M = T.M, B = T.B, W = T.W;
  unsigned long long rng = 42ull * 0x9E3779B97F4A7C15ull + (unsigned long long)(t + 1) * 0xBF58476D1CE4E5B9ull;
  int nseg = 0, i = Lr, k = 0;
  int seg_j = 0, seg_m = 0, seg_i = 0, seg_k = 0, have_m = 0;
  char st = 'C';
  int guard = 0;
  while (st != 'S' && guard++ < 8 * (Lr + M) + 64) {
    float pr[4];
    switch (st) {
      case 'C':
        if (i == 0) { st = 'S'; break; }
        pr[0] = scalbnf(__fmul_rn(sf[i - 1].xC, xf.tCC), ef[i - 1] - ef[i]);
        pr[1] = __fmul_rn(sf[i].xE, xf.tEC);
        if (md_choose(pr, 2, md_uniform(rng)) == 0) i--;
        else st = 'E';
        break;
      case 'J':
        if (i == 0) { st = 'S'; break; }
        pr[0] = scalbnf(__fmul_rn(sf[i - 1].xJ, xf.tJJ), ef[i - 1] - ef[i]);
        pr[1] = __fmul_rn(sf[i].xE, xf.tEJ);
        if (md_choose(pr, 2, md_uniform(rng)) == 0) i--;
        else st = 'E';
        break;
      case 'E': {  // all M_k, D_k of row i, k ascending (k = l*B + j + 1 lives in cell j*32 + l), M before D
        const float *Mi = FM + (size_t)i * W, *Di = FD + (size_t)i * W;
        float tot = 0.0f;
        for (int l = 0, kk = 1; l < 32 && kk <= M; l++)
          for (int j = 0; j < B && kk <= M; j++, kk++) { tot = __fadd_rn(tot, Mi[j * 32 + l]); tot = __fadd_rn(tot, Di[j * 32 + l]); }
        const float r = __fmul_rn(md_uniform(rng), tot);
        float cum = 0.0f;
        int ck = M, found = 0;
        char cs = 'M';
        for (int l = 0, kk = 1; l < 32 && kk <= M && !found; l++)
          for (int j = 0; j < B && kk <= M; j++, kk++) {
            const float vm = Mi[j * 32 + l], vd = Di[j * 32 + l];
            if (vm > 0.0f) { ck = kk; cs = 'M'; }
            cum = __fadd_rn(cum, vm);
            if (r < cum) { ck = kk; cs = 'M'; found = 1; break; }
            if (vd > 0.0f) { ck = kk; cs = 'D'; }
            cum = __fadd_rn(cum, vd);
            if (r < cum) { ck = kk; cs = 'D'; found = 1; break; }
          }
        k = ck;
        st = cs;
        have_m = 0;
      } break;
      case 'M': {
        if (!have_m) { seg_j = i; seg_m = k; have_m = 1; }
        seg_i = i; seg_k = k;
        const float4 A = T.trA[cell_of_k(k, B)];
        const size_t rb = (size_t)(i - 1) * W;
        const int cp = k >= 2 ? cell_of_k(k - 1, B) : 0;
        pr[0] = __fmul_rn(sf[i - 1].xB, A.x);
        pr[1] = k >= 2 ? __fmul_rn(FM[rb + cp], A.y) : 0.0f;
        pr[2] = k >= 2 ? __fmul_rn(FI[rb + cp], A.z) : 0.0f;
        pr[3] = k >= 2 ? __fmul_rn(FD[rb + cp], A.w) : 0.0f;
        const int c = md_choose(pr, 4, md_uniform(rng));
        i--;
        if (c == 0) st = 'B';
        else { k--; st = c == 1 ? 'M' : c == 2 ? 'I' : 'D'; }
      } break;
      case 'D': {
        const float4 Bq = T.trB[cell_of_k(k, B)];
        const int cp = k >= 2 ? cell_of_k(k - 1, B) : 0;
        pr[0] = k >= 2 ? __fmul_rn(FM[(size_t)i * W + cp], Bq.x) : 0.0f;
        pr[1] = k >= 2 ? __fmul_rn(FD[(size_t)i * W + cp], Bq.y) : 0.0f;
        const int c = md_choose(pr, 2, md_uniform(rng));
        k--;
        st = c == 0 ? 'M' : 'D';
      } break;
      case 'I': {
        const int c0 = cell_of_k(k, B);
        const float4 Bq = T.trB[c0];
        pr[0] = __fmul_rn(FM[(size_t)(i - 1) * W + c0], Bq.z);
        pr[1] = __fmul_rn(FI[(size_t)(i - 1) * W + c0], Bq.w);
        const int c = md_choose(pr, 2, md_uniform(rng));
        i--;
        st = c == 0 ? 'M' : 'I';
      } break;
      case 'B': {
        if (have_m && nseg < PA_MAXSEG) { seg4[4 * nseg] = seg_i; seg4[4 * nseg + 1] = seg_j; seg4[4 * nseg + 2] = seg_k; seg4[4 * nseg + 3] = seg_m; nseg++; }
        else if (have_m) atomicOr(status, 2);
        have_m = 0;
        pr[0] = __fmul_rn(sf[i].xN, xf.tNB);
        pr[1] = __fmul_rn(sf[i].xJ, xf.tJB);
        st = md_choose(pr, 2, md_uniform(rng)) == 0 ? 'N' : 'J';
      } break;
      case 'N':
        if (i == 0) st = 'S';
        else i--;
        break;
      default: st = 'S';
    }
  }
  return nseg;
}

__device__ __forceinline__ bool md_linked(const int *a, const int *b) {  // {t, i, j, k, m}
  int nov = min(a[2], b[2]) - max(a[1], b[1]) + 1;
  int n = min(a[2] - a[1] + 1, b[2] - b[1] + 1);
  if (__fdiv_rn((float)nov, (float)n) < 0.8f) return false;
  nov = min(a[4], b[4]) - max(a[3], b[3]) + 1;
  n = min(a[4] - a[3] + 1, b[4] - b[3] + 1);
  if (__fdiv_rn((float)nov, (float)n) < 0.8f) return false;
  if (abs((a[3] - a[1]) - (b[3] - b[1])) > 4) return false;
  if (abs((a[4] - a[2]) - (b[4] - b[2])) > 4) return false;
  return true;
}

// Envelopes of the multi-domain region ireg..jreg -> envl[2e], envl[2e+1] (target coordinates), ordered by
// start then end. Whole warp. scratch: segraw 4*NS*MAXSEG | segc 5*NS*MAXSEG | lab NS*MAXSEG | nper NS; cnt_i/cnt_j: Lr+2 ints.
__device__ int resolve_multidomain(const FwdTabs &T, const uint8_t *sub, int Lr, int Lfull, int ireg, float *rows, float *FM,
                                   float *FI, float *FD, Specials *sf, int *ef, int *segscratch, int *cnt_i, int *cnt_j,
                                   int *envl, int lane, int *status) {
  const XF xf = xf_config(Lfull, 1);
  int Ze;
  float Zm;
  forward_full(T, sub, Lr, xf, rows, FM, FI, FD, sf, ef, lane, &Ze, &Zm);
  __syncwarp();
  int *segraw = segscratch, *segc = segraw + 4 * PA_NSAMPLES * PA_MAXSEG, *lab = segc + 5 * PA_NSAMPLES * PA_MAXSEG,
      *nper = lab + PA_NSAMPLES * PA_MAXSEG;
  for (int t = lane; t < PA_NSAMPLES; t += 32)
    nper[t] = md_sample_trace(T, FM, FI, FD, sf, ef, Lr, xf, t, segraw + 4 * t * PA_MAXSEG, status);
  __syncwarp();
  int nseg = 0;
  if (lane == 0) {  // compact, trace by trace
    for (int t = 0; t < PA_NSAMPLES; t++)
      for (int q = 0; q < nper[t]; q++) {
        const int *r = segraw + 4 * (t * PA_MAXSEG + q);
        int *d = segc + 5 * nseg;
        d[0] = t; d[1] = r[0]; d[2] = r[1]; d[3] = r[2]; d[4] = r[3];
        lab[nseg] = nseg;
        nseg++;
      }
  }
  nseg = __shfl_sync(PA_FULL, nseg, 0);
  __syncwarp();
  // single linkage by label propagation: label = smallest member index of the connected component
  for (;;) {
    int changed = 0;
    for (int a = lane; a < nseg; a += 32) {
      int la = ((volatile int *)lab)[a];
      for (int b = 0; b < nseg; b++) {
        const int lb = ((volatile int *)lab)[b];
        if (lb < la && md_linked(segc + 5 * a, segc + 5 * b)) { la = lb; changed = 1; }
      }
      ((volatile int *)lab)[a] = la;
    }
    __syncwarp();
    if (!__any_sync(PA_FULL, changed)) break;
  }
  int nenv = 0;
  if (lane == 0) {
    for (int z = 0; z <= Lr; z++) { cnt_i[z] = 0; cnt_j[z] = 0; }
    for (int c = 0; c < nseg; c++) {
      if (lab[c] != c) continue;
      int ninc = 0, lastt = -1, imin = Lr + 1, imax = 0, jmin = Lr + 1, jmax = 0;
      for (int a = c; a < nseg; a++) {
        if (lab[a] != c) continue;
        const int *s = segc + 5 * a;
        if (s[0] != lastt) { ninc++; lastt = s[0]; }
        cnt_i[s[1]]++;
        cnt_j[s[2]]++;
        imin = min(imin, s[1]); imax = max(imax, s[1]);
        jmin = min(jmin, s[2]); jmax = max(jmax, s[2]);
      }
      if (__fdiv_rn((float)ninc, (float)PA_NSAMPLES) >= 0.25f) {
        int bi, bj, arg = imin;
        for (bi = imin; bi <= imax; bi++) {
          if (cnt_i[bi] > cnt_i[arg]) arg = bi;
          if (__fdiv_rn((float)cnt_i[bi], (float)ninc) >= 0.02f) break;
        }
        if (bi > imax) bi = arg;
        arg = jmax;
        for (bj = jmax; bj >= jmin; bj--) {
          if (cnt_j[bj] > cnt_j[arg]) arg = bj;
          if (__fdiv_rn((float)cnt_j[bj], (float)ninc) >= 0.02f) break;
        }
        if (bj < jmin) bj = arg;
        if (nenv < PA_MAXENV && bj >= bi) { envl[2 * nenv] = bi + ireg - 1; envl[2 * nenv + 1] = bj + ireg - 1; nenv++; }
        else if (bj >= bi) atomicOr(status, 2);
      }
      for (int a = c; a < nseg; a++)
        if (lab[a] == c) { cnt_i[segc[5 * a + 1]] = 0; cnt_j[segc[5 * a + 2]] = 0; }
    }
    for (int a = 1; a < nenv; a++)
      for (int b = a; b > 0 && (envl[2 * b] < envl[2 * b - 2] || (envl[2 * b] == envl[2 * b - 2] && envl[2 * b + 1] < envl[2 * b - 1])); b--) {
        const int ti = envl[2 * b], tj = envl[2 * b + 1];
        envl[2 * b] = envl[2 * b - 2]; envl[2 * b + 1] = envl[2 * b - 1];
        envl[2 * b - 2] = ti; envl[2 * b - 1] = tj;
      }
  }
  nenv = __shfl_sync(PA_FULL, nenv, 0);
  __syncwarp();
  return nenv;
}

static __global__ void __launch_bounds__(128, PA_DOM_MINB) domain_kernel(DomArgs a) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float *slab = a.scratch + (size_t)gw * a.slab_floats;
  const int maxL1 = a.maxL + 1;
  const SlabLayout sl = slab_layout(a.maxL, a.maxW, a.maxM);
  float *rows = slab;                                  // 8 * maxW
  Specials *sf = reinterpret_cast<Specials *>(slab + sl.sf);
  Specials *sb = reinterpret_cast<Specials *>(slab + sl.sb);
  int *ef = reinterpret_cast<int *>(slab + sl.ef), *eb = reinterpret_cast<int *>(slab + sl.eb);
  float *btot = slab + sl.lin, *etot = btot + maxL1, *mocc = etot + maxL1;
  float *ppN = mocc + maxL1, *ppJ = ppN + maxL1, *ppC = ppJ + maxL1;
  float *oE = ppC + maxL1, *oN = oE + maxL1, *oJ = oN + maxL1, *oB = oJ + maxL1, *oC = oB + maxL1;
  int *reg = reinterpret_cast<int *>(slab + sl.reg);
  int *trk = reinterpret_cast<int *>(slab + sl.trk), *tri = reinterpret_cast<int *>(slab + sl.tri);
  char *trs = reinterpret_cast<char *>(slab + sl.trs);
  int *segscratch = reinterpret_cast<int *>(slab + sl.seg);
  int *envl = segscratch + 10 * PA_NSAMPLES * PA_MAXSEG + PA_NSAMPLES;
  float *FM = slab + sl.mat;                           // 6 matrices of maxL1 * maxW
  const size_t msz = sl.msz;
  float *FI = FM + msz, *FD = FI + msz, *BM = FD + msz, *BI = BM + msz, *BD = BI + msz;
  const char *syms = a.abc == 0 ? "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~" : "ACGT-RYMKSWHBVDN*~";
  const int K = a.abc == 0 ? 20 : 4;

  for (;;) {
    unsigned long long w = 0;
    if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
    w = __shfl_sync(PA_FULL, w, 0);
    if (w >= a.n_in) break;
    const Pass4 r = a.in[w];
    const DevHmm &hm = a.hmms[r.h];
    const FwdTabs T = fwd_tabs(hm, a.tab);
    const int L = (int)a.tg.len[r.t], W = T.W, B = T.B, M = T.M;
    const uint8_t *p = a.tg.dsq + a.tg.off[r.t];
    // ---- multihit parsers -> specials
    float mant;
    const int Ze0 = fwd_parser(T, p, L, L, 1, rows, lane, &mant, sf, ef);
    const float Zm0 = mant;
    for (int c = lane; c < 8 * W; c += 32) rows[c] = 0.0f;
    backward_pass(T, p, L, L, 1, rows, nullptr, nullptr, nullptr, sb, eb, lane);
    __syncwarp();
    // ---- domain decoding + region scan (lane 0, serial in i)
    int nreg = 0;
    if (lane == 0) {
      const XF xf = xf_config(L, 1);
      btot[0] = etot[0] = mocc[0] = 0.0f;
      for (int i = 1; i <= L; i++) {
        const float pb = __fdiv_rn(scalbnf(__fmul_rn(sf[i - 1].xB, sb[i - 1].xB), ef[i - 1] + eb[i - 1] - Ze0), Zm0);
        btot[i] = __fadd_rn(btot[i - 1], pb);
        const float pe = __fdiv_rn(scalbnf(__fmul_rn(sf[i].xE, sb[i].xE), ef[i] + eb[i] - Ze0), Zm0);
        etot[i] = __fadd_rn(etot[i - 1], pe);
        const int ex = ef[i - 1] + eb[i] - Ze0;
        const float n = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xN, sb[i].xN), xf.tNN), ex), Zm0);
        const float jj = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xJ, sb[i].xJ), xf.tJJ), ex), Zm0);
        const float cc = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xC, sb[i].xC), xf.tCC), ex), Zm0);
        mocc[i] = __fsub_rn(1.0f, __fadd_rn(__fadd_rn(n, jj), cc));
      }
      const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
      int i = -1, trig = 0;
      for (int j = 1; j <= L; j++) {
        if (!trig) {
          if (__fsub_rn(mocc[j], __fsub_rn(btot[j], btot[j - 1])) < rt2) i = j;
          else if (i == -1) i = j;
          if (mocc[j] >= rt1) trig = 1;
        } else if (__fsub_rn(mocc[j], __fsub_rn(etot[j], etot[j - 1])) < rt2) {
          float mx = -1.0f;
          for (int z = i; z <= j; z++) {
            const float aa = __fsub_rn(etot[z], etot[i - 1]), bb = __fsub_rn(btot[j], btot[z - 1]);
            mx = fmaxf(mx, fminf(aa, bb));
          }
          if (nreg < PA_MAXREG) { reg[3 * nreg] = i; reg[3 * nreg + 1] = j; reg[3 * nreg + 2] = mx >= rt3; nreg++; }
          else atomicOr(a.status, 2);
          i = -1;
          trig = 0;
        }
      }
    }
    nreg = __shfl_sync(PA_FULL, nreg, 0);
    __syncwarp();
    int ndom = 0;
    for (int rg = 0; rg < nreg; rg++) {
     const int ireg = reg[3 * rg], jreg = reg[3 * rg + 1], multi = reg[3 * rg + 2];
     int nenv = 1;
     if (multi) {
       nenv = resolve_multidomain(T, p + ireg - 1, jreg - ireg + 1, L, ireg, rows, FM, FI, FD, sf, ef, segscratch,
                                  reinterpret_cast<int *>(ppN), reinterpret_cast<int *>(ppJ), envl, lane, a.status);
     } else if (lane == 0) { envl[0] = ireg; envl[1] = jreg; }
     __syncwarp();
     for (int en = 0; en < nenv; en++) {
      const int ienv = envl[2 * en], jenv = envl[2 * en + 1];
      const int Ld = jenv - ienv + 1;
      const uint8_t *sub = p + ienv - 1;  // sub[0..Ld-1]
      const XF xf = xf_config(L, 0);
      // ---- unihit Forward, full matrix
      int Ze = 0;
      float Zm;
      forward_full(T, sub, Ld, xf, rows, FM, FI, FD, sf, ef, lane, &Ze, &Zm);
      // ---- unihit Backward, full matrix
      for (int c = lane; c < 8 * W; c += 32) rows[c] = 0.0f;
      backward_pass(T, sub, Ld, L, 0, rows, BM, BI, BD, sb, eb, lane);
      __syncwarp();
      // ---- posterior decoding (B* <- pp), specials
      for (int i = 1; i <= Ld; i++) {
        const int ex = ef[i] + eb[i] - Ze;
        float *fM = FM + (size_t)i * W, *fI = FI + (size_t)i * W, *bM = BM + (size_t)i * W, *bI = BI + (size_t)i * W;
        for (int j = 0; j < B; j++) {
          const int c = j * 32 + lane;
          bM[c] = __fdiv_rn(scalbnf(__fmul_rn(fM[c], bM[c]), ex), Zm);
          bI[c] = __fdiv_rn(scalbnf(__fmul_rn(fI[c], bI[c]), ex), Zm);
        }
      }
      for (int i = 1 + lane; i <= Ld; i += 32) {
        const int ex2 = ef[i - 1] + eb[i] - Ze;
        ppN[i] = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xN, sb[i].xN), xf.tNN), ex2), Zm);
        ppJ[i] = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xJ, sb[i].xJ), xf.tJJ), ex2), Zm);
        ppC[i] = __fdiv_rn(scalbnf(__fmul_rn(__fmul_rn(sf[i - 1].xC, sb[i].xC), xf.tCC), ex2), Zm);
      }
      __syncwarp();
      // ---- null2 by expectation (before OA overwrites nothing it needs; uses pp in B*)
      float null2[32];
#pragma unroll
      for (int x = 0; x < 32; x++) null2[x] = 1.0f;
      if (a.do_null2) {
        float *eM = rows, *eI = rows + W;  // rows were zeroed before backward_pass; re-zero own cells
        for (int j = 0; j < B; j++) { eM[j * 32 + lane] = 0.0f; eI[j * 32 + lane] = 0.0f; }
        for (int i = 1; i <= Ld; i++) {
          const float *pM = BM + (size_t)i * W, *pI = BI + (size_t)i * W;
          for (int j = 0; j < B; j++) {
            const int c = j * 32 + lane;
            eM[c] = __fadd_rn(eM[c], pM[c]);
            eI[c] = __fadd_rn(eI[c], pI[c]);
          }
        }
        float sN = 0.0f, sJ = 0.0f, sC = 0.0f;
        for (int i = 1; i <= Ld; i++) { sN = __fadd_rn(sN, ppN[i]); sJ = __fadd_rn(sJ, ppJ[i]); sC = __fadd_rn(sC, ppC[i]); }
        const float norm = __fdiv_rn(1.0f, (float)Ld);
        const int nv = min(B, max(0, M - lane * B));
        for (int j = 0; j < nv; j++) {
          const int c = j * 32 + lane;
          eM[c] = __fmul_rn(eM[c], norm);
          eI[c] = __fmul_rn(eI[c], norm);
        }
        const float xfactor = __fadd_rn(__fadd_rn(__fmul_rn(sN, norm), __fmul_rn(sC, norm)), __fmul_rn(sJ, norm));
        for (int x = 0; x < K; x++) {
          const float *rf = T.em + (size_t)x * W;
          float s = 0.0f;
          for (int j = 0; j < nv; j++) {
            const int c = j * 32 + lane;
            s = __fadd_rn(s, __fmul_rn(eM[c], rf[c]));
            s = __fadd_rn(s, eI[c]);
          }
          null2[x] = __fadd_rn(lane_sum(s), xfactor);
        }
      }
      // ---- optimal accuracy DP (F* <- OA values), row 0 = -inf
      const float NEGINF = __int_as_float(0xff800000);
      for (int j = 0; j < B; j++) { FM[j * 32 + lane] = NEGINF; FI[j * 32 + lane] = NEGINF; FD[j * 32 + lane] = NEGINF; }
      if (lane == 0) { oN[0] = 0.0f; oB[0] = 0.0f; oE[0] = oJ[0] = oC[0] = NEGINF; }
      __syncwarp();
      {
        const int nv = min(B, max(0, M - lane * B));
        const int ilast = (B - 1) * 32 + lane;
        float poB = 0.0f, poJ = NEGINF, poC = NEGINF, poN = 0.0f;
        for (int i = 1; i <= Ld; i++) {
          float *Mp = FM + (size_t)(i - 1) * W, *Ip = FI + (size_t)(i - 1) * W, *Dp = FD + (size_t)(i - 1) * W;
          float *Mc = FM + (size_t)i * W, *Ic = FI + (size_t)i * W, *Dc = FD + (size_t)i * W;
          const float *pM = BM + (size_t)i * W, *pI = BI + (size_t)i * W;
          float mL = __shfl_up_sync(PA_FULL, Mp[ilast], 1), iL = __shfl_up_sync(PA_FULL, Ip[ilast], 1),
                dL = __shfl_up_sync(PA_FULL, Dp[ilast], 1);
          if (lane == 0) mL = iL = dL = 0.0f;
          float rmax = 0.0f;
          for (int j = 0; j < nv; j++) {
            const int c = j * 32 + lane;
            const float4 A = T.trA[c], Bq = T.trB[c];
            const float pm = j ? Mp[c - 32] : mL, pi = j ? Ip[c - 32] : iL, pd = j ? Dp[c - 32] : dL;
            float sv = A.x > 0.0f ? poB : 0.0f, v;
            v = A.y > 0.0f ? pm : 0.0f; if (v > sv) sv = v;
            v = A.z > 0.0f ? pi : 0.0f; if (v > sv) sv = v;
            v = A.w > 0.0f ? pd : 0.0f; if (v > sv) sv = v;
            const float mc = __fadd_rn(sv, pM[c]);
            const float ia = Bq.z > 0.0f ? Mp[c] : 0.0f, ib = Bq.w > 0.0f ? Ip[c] : 0.0f;
            Mc[c] = mc;
            Ic[c] = __fadd_rn(ia > ib ? ia : ib, pI[c]);
            if (mc > rmax) rmax = mc;
          }
          // D: D[k] = max(MD(k-1)? M[k-1]:0, DD(k-1)? D[k-1]:0) -- max is exact in any order
          float mcL = __shfl_up_sync(PA_FULL, Mc[ilast], 1);
          if (lane == 0) mcL = 0.0f;
          float LA = 0.0f;
          int LPok = 1;
          for (int j = 0; j < nv; j++) {
            const int c = j * 32 + lane;
            const float4 Bq = T.trB[c];
            const float ml = j ? Mc[c - 32] : mcL;
            const float aa = Bq.x > 0.0f ? ml : 0.0f;
            const int ddok = Bq.y > 0.0f;
            const float bb = (j && ddok) ? LA : 0.0f;
            LA = aa > bb ? aa : bb;
            LPok = (j ? LPok : 1) & ddok;
            Dc[c] = LA;
            BD[(size_t)i * W + c] = LPok ? 1.0f : 0.0f;  // reuse the (unused) backward-D row as the pass-through flag
          }
          if (nv == 0) { LA = 0.0f; LPok = 1; }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float a2 = __shfl_up_sync(PA_FULL, LA, o);
            const int p2 = __shfl_up_sync(PA_FULL, LPok, o);
            if (lane >= o) {
              const float cand = LPok ? a2 : 0.0f;
              if (cand > LA) LA = cand;
              LPok &= p2;
            }
          }
          float din = __shfl_up_sync(PA_FULL, LA, 1);
          if (lane == 0) din = 0.0f;
          for (int j = 0; j < nv; j++) {
            const int c = j * 32 + lane;
            const float cand = BD[(size_t)i * W + c] != 0.0f ? din : 0.0f;
            float d = Dc[c];
            if (cand > d) d = cand;
            Dc[c] = d;
            if (d > rmax) rmax = d;
          }
          const float xE = warp_fmax(rmax);
          float t1, t2;
          t1 = (xf.tJJ == 0.0f) ? 0.0f : __fadd_rn(poJ, ppJ[i]);
          t2 = (xf.tEJ == 0.0f) ? 0.0f : xE;
          const float cJ = (xf.tJJ == 0.0f && xf.tEJ == 0.0f) ? NEGINF : (t1 > t2 ? t1 : t2);
          t1 = (xf.tCC == 0.0f) ? 0.0f : __fadd_rn(poC, ppC[i]);
          t2 = (xf.tEC == 0.0f) ? 0.0f : xE;
          const float cC = t1 > t2 ? t1 : t2;
          const float cN = (xf.tNN == 0.0f) ? 0.0f : __fadd_rn(poN, ppN[i]);
          t1 = (xf.tNB == 0.0f) ? 0.0f : cN;
          t2 = (xf.tJB == 0.0f) ? 0.0f : cJ;
          const float cB = t1 > t2 ? t1 : t2;
          if (lane == 0) { oE[i] = xE; oJ[i] = cJ; oC[i] = cC; oN[i] = cN; oB[i] = cB; }
          poB = cB; poJ = cJ; poC = cC; poN = cN;
        }
      }
      __syncwarp();
      // ---- traceback on lane 0
      int ali_len = 0, hmmfrom = 0, hmmto = 0, iali = 0, jali = 0, ok = 0, z1 = -1, z2 = -1, ntr = 0;
      if (lane == 0) {
        int i = Ld, k = 0, fail = 0, guard = 0;
        char sprv = 'C', scur = 'C';
        const int cap = a.maxL + a.maxM + 16;
        trs[ntr] = 'T'; trk[ntr] = 0; tri[ntr] = 0; ntr++;
        trs[ntr] = 'C'; trk[ntr] = 0; tri[ntr] = 0; ntr++;
#define OKT(kk, which) (((kk) >= 1) && (which) > 0.0f)
        while (sprv != 'S' && !fail) {
          switch (sprv) {
            case 'M': {
              const float4 A = T.trA[cell_of_k(k, B)];
              float path[4];
              const size_t rb = (size_t)(i - 1) * W;
              const int cp = k >= 2 ? cell_of_k(k - 1, B) : 0;
              path[0] = (k >= 2 && A.y > 0.0f) ? FM[rb + cp] : NEGINF;
              path[1] = (k >= 2 && A.z > 0.0f) ? FI[rb + cp] : NEGINF;
              path[2] = (k >= 2 && A.w > 0.0f) ? FD[rb + cp] : NEGINF;
              path[3] = (A.x > 0.0f) ? oB[i - 1] : NEGINF;
              int best = 0;
              for (int z = 1; z < 4; z++) if (path[z] > path[best]) best = z;
              scur = "MIDB"[best];
              k--; i--;
            } break;
            case 'D': {
              const float4 Bq = T.trB[cell_of_k(k, B)];
              const int cp = k >= 2 ? cell_of_k(k - 1, B) : 0;
              const float aa = (k >= 2 && Bq.x > 0.0f) ? FM[(size_t)i * W + cp] : NEGINF;
              const float bb = (k >= 2 && Bq.y > 0.0f) ? FD[(size_t)i * W + cp] : NEGINF;
              scur = (aa >= bb) ? 'M' : 'D';
              k--;
            } break;
            case 'I': {
              const int c = cell_of_k(k, B);
              const float4 Bq = T.trB[c];
              const float aa = (Bq.z > 0.0f) ? FM[(size_t)(i - 1) * W + c] : NEGINF;
              const float bb = (Bq.w > 0.0f) ? FI[(size_t)(i - 1) * W + c] : NEGINF;
              scur = (aa >= bb) ? 'M' : 'I';
              i--;
            } break;
            case 'N': scur = (i == 0) ? 'S' : 'N'; break;
            case 'C': {
              if (i == 0) { fail = 1; scur = 'S'; break; }
              const float aa = (xf.tCC == 0.0f) ? NEGINF : __fadd_rn(oC[i - 1], ppC[i]);
              const float bb = (xf.tEC == 0.0f) ? NEGINF : oE[i];
              scur = (aa > bb) ? 'C' : 'E';
            } break;
            case 'J': {
              const float aa = (xf.tJJ == 0.0f) ? NEGINF : __fadd_rn(oJ[i - 1], ppJ[i]);
              const float bb = (xf.tEJ == 0.0f) ? NEGINF : oE[i];
              scur = (aa > bb) ? 'J' : 'E';
            } break;
            case 'E': {
              float mx = NEGINF;
              int kmax = 0;
              char smax = 'M';
              for (int kk = 1; kk <= M; kk++) {
                const int c = cell_of_k(kk, B);
                const float vm = FM[(size_t)i * W + c], vd = FD[(size_t)i * W + c];
                if (vm > mx) { mx = vm; kmax = kk; smax = 'M'; }
                if (vd > mx) { mx = vd; kmax = kk; smax = 'D'; }
              }
              k = kmax;
              scur = smax;
            } break;
            case 'B': {
              const float aa = (xf.tNB == 0.0f) ? NEGINF : oN[i];
              const float bb = (xf.tJB == 0.0f) ? NEGINF : oJ[i];
              scur = (aa > bb) ? 'N' : 'J';
            } break;
            default: fail = 1; scur = 'S';
          }
          if (scur == 'S' || ntr >= cap) { if (ntr >= cap) fail = 1; break; }
          if (scur == 'M' || scur == 'I' || scur == 'D') { trs[ntr] = scur; trk[ntr] = k; tri[ntr] = (scur == 'D') ? 0 : i; ntr++; }
          else if ((scur == 'N' || scur == 'J' || scur == 'C') && scur == sprv) { trs[ntr] = scur; trk[ntr] = 0; tri[ntr] = i; ntr++; i--; }
          else { trs[ntr] = scur; trk[ntr] = 0; tri[ntr] = 0; ntr++; }
          sprv = scur;
          if (++guard > 4 * (Ld + M) + 16) fail = 1;
        }
#undef OKT
        if (!fail) {
          // trace is stored backwards: first M in sequence order = last M pushed
          for (int z = ntr - 1; z >= 0; z--) if (trs[z] == 'M') { z1 = z; break; }
          for (int z = 0; z < ntr; z++) if (trs[z] == 'M') { z2 = z; break; }
          if (z1 >= 0) {
            ok = 1;
            hmmfrom = trk[z1]; hmmto = trk[z2];
            iali = tri[z1] + ienv - 1; jali = tri[z2] + ienv - 1;
            for (int z = z1; z >= z2; z--) if (trs[z] == 'M' || trs[z] == 'I' || trs[z] == 'D') ali_len++;
          }
        }
      }
      ok = __shfl_sync(PA_FULL, ok, 0);
      if (!ok) continue;
      // ---- emit the domain record
      unsigned long long ridx = 0, poff = 0, roff = 0;
      if (lane == 0) {
        ridx = atomicAdd(a.rec_count, 1ull);
        poff = atomicAdd(a.pool_count, (unsigned long long)(2 * (ali_len + 1)));  // aligned line, then its PP line
        roff = atomicAdd(a.res_count, (unsigned long long)Ld);
        if (ridx >= a.rec_cap || poff + 2 * (ali_len + 1) > a.pool_cap || roff + Ld > a.res_cap) {
          atomicOr(a.status, 1);
        } else {
          const char *cons = a.cons + a.cons_off[r.h];
          int n = 0;
          for (int z = z1; z >= z2; z--) {
            const char st = trs[z];
            char mch, tch, pch;
            // posterior probability of the aligned state (B* holds pp since the decoding step): HMMER's PP line
            if (st == 'M') { mch = cons[trk[z]]; tch = syms[sub[tri[z] - 1]]; pch = pp_char(BM[(size_t)tri[z] * W + cell_of_k(trk[z], B)]); }
            else if (st == 'I') { mch = '.'; tch = syms[sub[tri[z] - 1]] | 0x20; pch = pp_char(BI[(size_t)tri[z] * W + cell_of_k(trk[z], B)]); }
            else if (st == 'D') { mch = cons[trk[z]]; tch = '-'; pch = '.'; }
            else continue;
            a.model[poff + n] = mch;
            a.aligned[poff + n] = tch;
            a.aligned[poff + ali_len + 1 + n] = pch;
            n++;
          }
          a.model[poff + n] = 0;
          a.aligned[poff + n] = 0;
          a.aligned[poff + 2 * ali_len + 1] = 0;
          for (int q = 0; q < Ld; q++) a.res[roff + q] = sub[q];
          DomRec d;
          d.t = r.t; d.h = r.h; d.multi = (uint16_t)multi; d.surv = (int)w; d.dom_idx = ndom;
          d.ienv = ienv; d.jenv = jenv; d.iali = iali; d.jali = jali; d.hmmfrom = hmmfrom; d.hmmto = hmmto;
          d.env_exp = Ze; d.env_mant = Zm; d.oasc = oC[Ld];
          d.ali_off = (int)poff; d.ali_len = ali_len; d.res_off = (int)roff;
          for (int x = 0; x < 32; x++) d.null2[x] = x < K ? null2[x] : 1.0f;
          a.recs[ridx] = d;
        }
      }
      ndom++;
      __syncwarp();
     }
    }
  }
}

}  // namespace pa
