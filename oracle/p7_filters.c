/*
 * p7_filters.c -- ORACLE (test infrastructure): MSV (8-bit), Viterbi (16-bit),
 * bias filter, Forward/Backward parsers in scaled probability space, plus
 * textbook log-space Forward/Viterbi for cross-checks.
 * Restates SURVEY.md Appendix A.4-A.6 (HMMER3 is invoked by the reference at
 * /root/reference/lib/aln.py:664-667 and is not vendored).
 */
#include "p7oracle.h"
#include "p7_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#pragma STDC FP_CONTRACT OFF

/* ---------------- MSV: unsigned saturating bytes ---------------- */
static inline int sat_addu8(int a, int b) { int s = a + b; return s > 255 ? 255 : s; }
static inline int sat_subu8(int a, int b) { int s = a - b; return s < 0 ? 0 : s; }

int orc_msv(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xJ_raw) {
  int M = p->M;
  uint8_t tjb = p7i_unbiased_byteify(p->scale_b, logf(3.0f / (float)(L + 3)));
  int tjbm = (uint8_t)((int8_t)tjb + (int8_t)p->tbm_b); /* HMMER adds as int8, wraps */
  int bias = p->bias_b, base = p->base_b, tec = p->tec_b;
  uint8_t *dp = (uint8_t *)calloc((size_t)M + 1, 1);
  int xJ = 0, xB = sat_subu8(base, tjbm);
  for (int i = 1; i <= L; i++) {
    const uint8_t *rsc = p->rbv + (size_t)dsq[i] * (M + 1);
    int xE = 0, mpv = 0; /* dp[i-1][0] = 0 (-inf) */
    for (int k = 1; k <= M; k++) {
      int sv = mpv > xB ? mpv : xB;
      sv = sat_addu8(sv, bias);
      sv = sat_subu8(sv, rsc[k]);
      if (sv > xE) xE = sv;
      mpv = dp[k];
      dp[k] = (uint8_t)sv;
    }
    if (sat_addu8(xE, bias) == 255) { /* overflow: score is +inf, passes */
      free(dp);
      *ret_sc = INFINITY;
      if (xJ_raw) *xJ_raw = -1;
      return 1;
    }
    xE = sat_subu8(xE, tec);
    if (xE > xJ) xJ = xE;
    xB = sat_subu8(base > xJ ? base : xJ, tjbm);
  }
  free(dp);
  if (xJ_raw) *xJ_raw = xJ;
  float sc = ((float)(xJ - tjb) - (float)base);
  sc /= p->scale_b;
  sc -= 3.0f;
  *ret_sc = sc;
  return 0;
}

/* ---------------- MSV, striped AVX2 (the way HMMER's own filter is vectorised) ----------------
 * Farrar striping over 32 unsigned-byte lanes; saturating adds/subs are the hardware's.  Same
 * recurrence as orc_msv above, hence the same xJ byte and score: integer arithmetic is order-free.
 * This is what bench.py times as the host-core baseline. */
#if defined(__x86_64__)
#include <immintrin.h>
int orc_have_avx2(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }

__attribute__((target("avx2"))) int orc_msv_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xJ_raw) {
  if (!p->rbv_s || !orc_have_avx2()) return -1;
  const int Q = p->Qb;
  uint8_t tjb = p7i_unbiased_byteify(p->scale_b, logf(3.0f / (float)(L + 3)));
  int tjbm = (uint8_t)((int8_t)tjb + (int8_t)p->tbm_b);
  int bias = p->bias_b, base = p->base_b, tec = p->tec_b;
  __m256i dpbuf[64];
  __m256i *dp = Q <= 64 ? dpbuf : (__m256i *)aligned_alloc(32, (size_t)Q * 32);
  for (int q = 0; q < Q; q++) dp[q] = _mm256_setzero_si256();
  const __m256i biasv = _mm256_set1_epi8((char)bias);
  int xJ = 0, xB = sat_subu8(base, tjbm);
  int overflow = 0;
  for (int i = 1; i <= L; i++) {
    const __m256i *rsc = (const __m256i *)(p->rbv_s + (size_t)dsq[i] * Q * 32);
    const __m256i xBv = _mm256_set1_epi8((char)xB);
    __m256i xEv = _mm256_setzero_si256();
    /* previous row's last vector, shifted up one byte lane (lane 0 gets 0 = -inf) */
    __m256i last = dp[Q - 1];
    __m256i mpv = _mm256_alignr_epi8(last, _mm256_permute2x128_si256(last, last, 0x08), 15);
    for (int q = 0; q < Q; q++) {
      __m256i sv = _mm256_max_epu8(mpv, xBv);
      sv = _mm256_adds_epu8(sv, biasv);
      sv = _mm256_subs_epu8(sv, _mm256_load_si256(rsc + q));
      xEv = _mm256_max_epu8(xEv, sv);
      mpv = dp[q];
      dp[q] = sv;
    }
    __m128i m = _mm_max_epu8(_mm256_castsi256_si128(xEv), _mm256_extracti128_si256(xEv, 1));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 8));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 4));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 2));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 1));
    int xE = _mm_cvtsi128_si32(m) & 0xff;
    if (sat_addu8(xE, bias) == 255) { overflow = 1; break; }
    xE = sat_subu8(xE, tec);
    if (xE > xJ) xJ = xE;
    xB = sat_subu8(base > xJ ? base : xJ, tjbm);
  }
  if (dp != dpbuf) free(dp);
  if (overflow) {
    *ret_sc = INFINITY;
    if (xJ_raw) *xJ_raw = -1;
    return 1;
  }
  if (xJ_raw) *xJ_raw = xJ;
  float sc = ((float)(xJ - tjb) - (float)base);
  sc /= p->scale_b;
  sc -= 3.0f;
  *ret_sc = sc;
  return 0;
}
#else
int orc_have_avx2(void) { return 0; }
int orc_msv_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xJ_raw) {
  (void)p; (void)dsq; (void)L; (void)ret_sc; (void)xJ_raw;
  return -1;
}
#endif

/* ---------------- Viterbi filter: signed saturating words ---------------- */
static inline int sat_adds16(int a, int b) {
  int s = a + b;
  if (s > 32767) return 32767;
  if (s < -32768) return -32768;
  return s;
}
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

int orc_vit(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xC_raw) {
  int M = p->M;
  const int16_t *tw = p->twv;
  float pmove = (2.0f + 1.0f) / ((float)L + 2.0f + 1.0f);
  int xw_move = p7i_wordify(p->scale_w, logf(pmove));       /* N->B, J->B, C->T */
  int xw_E = p7i_wordify(p->scale_w, -0.69314718055994529f); /* E->C, E->J (multihit) */
  int *mmx = (int *)malloc(sizeof(int) * (M + 1) * 6);
  int *imx = mmx + (M + 1), *dmx = imx + (M + 1), *mc = dmx + (M + 1), *ic = mc + (M + 1), *dc = ic + (M + 1);
  for (int k = 0; k <= M; k++) mmx[k] = imx[k] = dmx[k] = -32768;
  int xN = p->base_w, xB = sat_adds16(xN, xw_move), xJ = -32768, xC = -32768, xE;
  for (int i = 1; i <= L; i++) {
    const int16_t *rsc = p->rwv + (size_t)dsq[i] * (M + 1);
    xE = -32768;
    mc[0] = ic[0] = dc[0] = -32768;
    for (int k = 1; k <= M; k++) {
      int sv = sat_adds16(xB, tw[(k - 1) * 8 + ORC_BM]);
      sv = MAX2(sv, sat_adds16(mmx[k - 1], tw[(k - 1) * 8 + ORC_MM]));
      sv = MAX2(sv, sat_adds16(imx[k - 1], tw[(k - 1) * 8 + ORC_IM]));
      sv = MAX2(sv, sat_adds16(dmx[k - 1], tw[(k - 1) * 8 + ORC_DM]));
      sv = sat_adds16(sv, rsc[k]);
      mc[k] = sv;
      xE = MAX2(xE, sv);
      /* D: full (non-lazy) propagation; identical score to HMMER's lazy-F (A.5) */
      dc[k] = MAX2(sat_adds16(mc[k - 1], tw[(k - 1) * 8 + ORC_MD]), sat_adds16(dc[k - 1], tw[(k - 1) * 8 + ORC_DD]));
      ic[k] = MAX2(sat_adds16(mmx[k], tw[k * 8 + ORC_MI]), sat_adds16(imx[k], tw[k * 8 + ORC_II]));
    }
    if (xE >= 32767) {
      free(mmx);
      *ret_sc = INFINITY;
      if (xC_raw) *xC_raw = 32767;
      return 1;
    }
    /* NN = CC = JJ = 0 in the filter */
    xC = MAX2(xC, xE + xw_E);
    xJ = MAX2(xJ, xE + xw_E);
    xB = MAX2(xJ + xw_move, xN + xw_move);
    if (xC < -32768) xC = -32768;
    if (xJ < -32768) xJ = -32768;
    if (xB < -32768) xB = -32768;
    memcpy(mmx, mc, sizeof(int) * (M + 1));
    memcpy(imx, ic, sizeof(int) * (M + 1));
    memcpy(dmx, dc, sizeof(int) * (M + 1));
  }
  free(mmx);
  if (xC_raw) *xC_raw = xC;
  if (xC > -32768) {
    float sc = (float)xC + (float)xw_move - (float)p->base_w;
    sc /= p->scale_w;
    sc -= 3.0f;
    *ret_sc = sc;
  } else
    *ret_sc = -INFINITY;
  return 0;
}

/* ---------------- Viterbi filter, striped AVX2 (16 signed saturating words per vector) ----------------
 * The recurrences of orc_vit above on Farrar's striped layout.  M and I of a row only need the previous row; D(k) also
 * needs D(k-1) of the SAME row, which crosses vector lanes: a first sweep assumes nothing enters a lane from its left
 * neighbour, then the D->D path is pushed across the lane boundaries until no cell changes ("lazy F").  Saturating max-plus
 * arithmetic gives the same words as the serial recurrence whatever the evaluation order. */
#if defined(__x86_64__)
__attribute__((target("avx2"))) static inline __m256i shift_words(__m256i v) {
  /* lane j <- lane j-1, lane 0 <- -32768 */
  __m256i t = _mm256_permute2x128_si256(v, v, 0x08);
  __m256i r = _mm256_alignr_epi8(v, t, 14);
  return _mm256_or_si256(r, _mm256_set_epi16(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (short)0x8000));
}

__attribute__((target("avx2"))) int orc_vit_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xC_raw) {
  if (!p->rwv_s || !p->twv_s || !orc_have_avx2()) return -1;
  const int Q = p->Qw;
  float pmove = (2.0f + 1.0f) / ((float)L + 2.0f + 1.0f);
  int xw_move = p7i_wordify(p->scale_w, logf(pmove));
  int xw_E = p7i_wordify(p->scale_w, -0.69314718055994529f);
  const __m256i *tBM = (const __m256i *)p->twv_s, *tMM = tBM + Q, *tIM = tMM + Q, *tDM = tIM + Q, *tMI = tDM + Q, *tII = tMI + Q,
                *tMD = tII + Q, *tDD = tMD + Q;
  __m256i stackbuf[3 * 48];
  __m256i *mmx = Q <= 48 ? stackbuf : (__m256i *)aligned_alloc(32, (size_t)3 * Q * 32);
  __m256i *imx = mmx + Q, *dmx = imx + Q;
  const __m256i neg = _mm256_set1_epi16(-32768);
  for (int q = 0; q < 3 * Q; q++) mmx[q] = neg;
  int xN = p->base_w, xB = sat_adds16(xN, xw_move), xJ = -32768, xC = -32768;
  int overflow = 0;
  for (int i = 1; i <= L; i++) {
    const __m256i *rsc = (const __m256i *)(p->rwv_s + (size_t)dsq[i] * Q * 16);
    const __m256i xBv = _mm256_set1_epi16((short)xB);
    __m256i mpv = shift_words(mmx[Q - 1]), ipv = shift_words(imx[Q - 1]), dpv = shift_words(dmx[Q - 1]);
    __m256i xEv = neg;
    for (int q = 0; q < Q; q++) {
      __m256i sv = _mm256_adds_epi16(xBv, tBM[q]);
      sv = _mm256_max_epi16(sv, _mm256_adds_epi16(mpv, tMM[q]));
      sv = _mm256_max_epi16(sv, _mm256_adds_epi16(ipv, tIM[q]));
      sv = _mm256_max_epi16(sv, _mm256_adds_epi16(dpv, tDM[q]));
      sv = _mm256_adds_epi16(sv, rsc[q]);
      xEv = _mm256_max_epi16(xEv, sv);
      mpv = mmx[q];
      ipv = imx[q];
      dpv = dmx[q];
      imx[q] = _mm256_max_epi16(_mm256_adds_epi16(mpv, tMI[q]), _mm256_adds_epi16(ipv, tII[q]));
      mmx[q] = sv;
    }
    __m128i m = _mm_max_epi16(_mm256_castsi256_si128(xEv), _mm256_extracti128_si256(xEv, 1));
    m = _mm_max_epi16(m, _mm_srli_si128(m, 8));
    m = _mm_max_epi16(m, _mm_srli_si128(m, 4));
    m = _mm_max_epi16(m, _mm_srli_si128(m, 2));
    int xE = (int)(short)(_mm_cvtsi128_si32(m) & 0xffff);
    if (xE >= 32767) { overflow = 1; break; }
    /* D of this row: first sweep inside the lanes */
    __m256i mleft = shift_words(mmx[Q - 1]), dleft = neg;
    for (int q = 0; q < Q; q++) {
      __m256i d = _mm256_max_epi16(_mm256_adds_epi16(mleft, tMD[q]), _mm256_adds_epi16(dleft, tDD[q]));
      dmx[q] = d;
      mleft = mmx[q];
      dleft = d;
    }
    /* then across the lanes until nothing changes (at most 16 sweeps: one per lane) */
    for (int pass = 0; pass < 16; pass++) {
      dleft = shift_words(dmx[Q - 1]);
      int q;
      for (q = 0; q < Q; q++) {
        __m256i cand = _mm256_adds_epi16(dleft, tDD[q]);
        if (_mm256_movemask_epi8(_mm256_cmpgt_epi16(cand, dmx[q])) == 0) break;
        dmx[q] = _mm256_max_epi16(dmx[q], cand);
        dleft = dmx[q];
      }
      if (q < Q) break;
    }
    xC = MAX2(xC, xE + xw_E);
    xJ = MAX2(xJ, xE + xw_E);
    xB = MAX2(xJ + xw_move, xN + xw_move);
    if (xC < -32768) xC = -32768;
    if (xJ < -32768) xJ = -32768;
    if (xB < -32768) xB = -32768;
  }
  if (mmx != stackbuf) free(mmx);
  if (overflow) {
    *ret_sc = INFINITY;
    if (xC_raw) *xC_raw = 32767;
    return 1;
  }
  if (xC_raw) *xC_raw = xC;
  if (xC > -32768) {
    float sc = (float)xC + (float)xw_move - (float)p->base_w;
    sc /= p->scale_w;
    sc -= 3.0f;
    *ret_sc = sc;
  } else
    *ret_sc = -INFINITY;
  return 0;
}
#else
int orc_vit_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xC_raw) {
  (void)p; (void)dsq; (void)L; (void)ret_sc; (void)xC_raw;
  return -1;
}
#endif

/* ---------------- bias filter: 2-state HMM forward ---------------- */
float orc_bias_filtersc(const ORC_PROFILE *p, const uint8_t *dsq, int L) {
  int K = p->K, Kp = p->Kp;
  float eo1[32];
  p7i_bias_eo1(p, eo1);
  (void)K;
  (void)Kp;
  float L0 = 400.0f, L1 = (float)p->M / 8.0f;
  float t00 = L0 / (L0 + 1.0f), t01 = 1.0f / (L0 + 1.0f);
  float t10 = 1.0f / (L1 + 1.0f), t11 = L1 / (L1 + 1.0f);
  float d0 = 0.999f, d1 = 0.001f * eo1[dsq[1]]; /* state-0 odds are exactly 1 */
  long sexp = 0;
  for (int i = 1;; i++) {
    float mx = d0 > d1 ? d0 : d1;
    int e = ilogbf(mx);
    d0 = scalbnf(d0, -e);
    d1 = scalbnf(d1, -e);
    sexp += e;
    if (i == L) break;
    float n0 = d0 * t00 + d1 * t10;
    float n1 = (d0 * t01 + d1 * t11) * eo1[dsq[i + 1]];
    d0 = n0;
    d1 = n1;
  }
  float tot = d0 + d1;
  float hmmsc = (float)(log((double)tot) + (double)sexp * 0.69314718055994529);
  float p1 = (float)L / (float)(L + 1);
  return hmmsc + (float)L * logf(p1) + logf((float)(1.0 - (double)p1));
}

/* ---------------- canonical scaled Forward / Backward ---------------- */

P7I_XF p7i_xf_config(int Lcfg, int multihit) {
  P7I_XF x;
  float nj = multihit ? 1.0f : 0.0f;
  float pmove = (2.0f + nj) / ((float)Lcfg + 2.0f + nj);
  float ploop = 1.0f - pmove;
  x.tNN = x.tJJ = x.tCC = ploop;
  x.tNB = x.tJB = x.tCT = pmove;
  x.tEJ = multihit ? 0.5f : 0.0f;
  x.tEC = multihit ? 0.5f : 1.0f;
  return x;
}

float p7i_lane_sum(const float *lane) { /* butterfly over 32 lanes: offsets 16,8,4,2,1 */
  float a[ORC_NL], b[ORC_NL];
  memcpy(a, lane, sizeof(a));
  for (int off = ORC_NL / 2; off >= 1; off >>= 1) {
    for (int l = 0; l < ORC_NL; l++) b[l] = a[l] + a[l ^ off];
    memcpy(a, b, sizeof(a));
  }
  return a[0];
}

/* One forward row. prev/cur rows are arrays [0..M]; index 0 is the zero boundary. Returns xE. */
float p7i_fwd_row(const ORC_PROFILE *p, int x, const float *Mp, const float *Ip, const float *Dp, float xBp,
                  float *Mc, float *Ic, float *Dc, float *wA, float *wP) {
  int M = p->M, B = (M + ORC_NL - 1) / ORC_NL;
  const float *tf = p->tfv;
  const float *rf = p->rfv + (size_t)x * (M + 1);
  Mc[0] = Ic[0] = Dc[0] = 0.0f;
  for (int k = 1; k <= M; k++) {
    float t = xBp * tf[(k - 1) * 8 + ORC_BM];
    t = t + Mp[k - 1] * tf[(k - 1) * 8 + ORC_MM];
    t = t + Ip[k - 1] * tf[(k - 1) * 8 + ORC_IM];
    t = t + Dp[k - 1] * tf[(k - 1) * 8 + ORC_DM];
    Mc[k] = t * rf[k];
    Ic[k] = Mp[k] * tf[k * 8 + ORC_MI] + Ip[k] * tf[k * 8 + ORC_II];
  }
  /* D chain: blocked scan over 32 lanes */
  float LA[ORC_NL], LP[ORC_NL];
  for (int l = 0; l < ORC_NL; l++) {
    int first = l * B + 1, last = (l + 1) * B;
    if (last > M) last = M;
    if (first > M) {
      LA[l] = 0.0f;
      LP[l] = 1.0f;
      continue;
    }
    for (int k = first; k <= last; k++) {
      float a = Mc[k - 1] * tf[(k - 1) * 8 + ORC_MD];
      if (k == first) {
        wA[k] = a;
        wP[k] = tf[(k - 1) * 8 + ORC_DD];
      } else {
        wA[k] = a + wA[k - 1] * tf[(k - 1) * 8 + ORC_DD];
        wP[k] = wP[k - 1] * tf[(k - 1) * 8 + ORC_DD];
      }
    }
    LA[l] = wA[last];
    LP[l] = wP[last];
  }
  for (int off = 1; off < ORC_NL; off <<= 1)
    for (int l = ORC_NL - 1; l >= off; l--) {
      LA[l] = LA[l] + LP[l] * LA[l - off];
      LP[l] = LP[l] * LP[l - off];
    }
  float lane[ORC_NL];
  for (int l = 0; l < ORC_NL; l++) {
    int first = l * B + 1, last = (l + 1) * B;
    if (last > M) last = M;
    float cin = (l == 0) ? 0.0f : LA[l - 1];
    float s = 0.0f;
    for (int k = first; k <= last; k++) {
      Dc[k] = wA[k] + wP[k] * cin;
      s = s + (Mc[k] + Dc[k]);
    }
    lane[l] = s;
  }
  return p7i_lane_sum(lane);
}

/* Cells of one backward row given we[k] = M(i+1,k)*emission (zeros for row L), In = I(i+1,.) and xE(i). */
void p7i_bwd_cells(const ORC_PROFILE *p, const float *we, const float *In, float xE, float *Mc, float *Ic,
                   float *Dc, float *wA, float *wP) {
  int M = p->M, B = (M + ORC_NL - 1) / ORC_NL;
  const float *tf = p->tfv;
  /* D chain descending, blocked */
  float LA[ORC_NL], LP[ORC_NL];
  for (int l = 0; l < ORC_NL; l++) {
    int first = l * B + 1, last = (l + 1) * B;
    if (last > M) last = M;
    if (first > M) {
      LA[l] = 0.0f;
      LP[l] = 1.0f;
      continue;
    }
    for (int k = last; k >= first; k--) {
      float c = (k == M) ? xE : (we[k + 1] * tf[k * 8 + ORC_DM] + xE);
      float dd = (k == M) ? 0.0f : tf[k * 8 + ORC_DD];
      if (k == last) {
        wA[k] = c;
        wP[k] = dd;
      } else {
        wA[k] = c + wA[k + 1] * dd;
        wP[k] = wP[k + 1] * dd;
      }
    }
    LA[l] = wA[first];
    LP[l] = wP[first];
  }
  for (int off = 1; off < ORC_NL; off <<= 1)
    for (int l = 0; l + off < ORC_NL; l++) {
      LA[l] = LA[l] + LP[l] * LA[l + off];
      LP[l] = LP[l] * LP[l + off];
    }
  for (int l = 0; l < ORC_NL; l++) {
    int first = l * B + 1, last = (l + 1) * B;
    if (last > M) last = M;
    float cin = (l == ORC_NL - 1) ? 0.0f : LA[l + 1];
    for (int k = first; k <= last; k++) Dc[k] = wA[k] + wP[k] * cin;
  }
  Dc[M + 1] = 0.0f;
  for (int k = 1; k <= M; k++) {
    if (k == M) {
      Mc[k] = xE;
      Ic[k] = 0.0f;
    } else {
      float t = we[k + 1] * tf[k * 8 + ORC_MM] + In[k] * tf[k * 8 + ORC_MI];
      t = t + Dc[k + 1] * tf[k * 8 + ORC_MD];
      Mc[k] = t + xE;
      Ic[k] = we[k + 1] * tf[k * 8 + ORC_IM] + In[k] * tf[k * 8 + ORC_II];
    }
  }
  Mc[0] = Ic[0] = Dc[0] = 0.0f;
}

/* One backward row i given row i+1 (Mn,In) and residue x = dsq[i+1].
 * Computes specials of row i into *sp (using next-row specials *sn), then M,I,D of row i. */
void p7i_bwd_row(const ORC_PROFILE *p, int x, const float *Mn, const float *In, const P7I_XF *xf,
                 const P7I_SPECIALS *sn, P7I_SPECIALS *sp, float *Mc, float *Ic, float *Dc, float *we,
                 float *wA, float *wP) {
  int M = p->M, B = (M + ORC_NL - 1) / ORC_NL;
  const float *tf = p->tfv;
  const float *rf = p->rfv + (size_t)x * (M + 1);
  float lane[ORC_NL];
  for (int k = 1; k <= M; k++) we[k] = Mn[k] * rf[k];
  we[M + 1] = 0.0f;
  for (int l = 0; l < ORC_NL; l++) {
    int first = l * B + 1, last = (l + 1) * B;
    if (last > M) last = M;
    float s = 0.0f;
    for (int k = first; k <= last; k++) s = s + we[k] * tf[(k - 1) * 8 + ORC_BM];
    lane[l] = s;
  }
  sp->xB = p7i_lane_sum(lane);
  sp->xJ = sp->xB * xf->tJB + sn->xJ * xf->tJJ;
  sp->xC = sn->xC * xf->tCC;
  sp->xE = sp->xJ * xf->tEJ + sp->xC * xf->tEC;
  sp->xN = sp->xB * xf->tNB + sn->xN * xf->tNN;
  p7i_bwd_cells(p, we, In, sp->xE, Mc, Ic, Dc, wA, wP);
}

#define P7I_RESCALE_THRESH 8192.0f

/* Forward over dsq[1..L]. If mx != NULL stores full rows ((L+1)*(M+2) floats each for M,I,D);
 * sp[0..L] specials (may be NULL), sexp[0..L] cumulative exponents (may be NULL). */
float p7i_forward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, P7I_MX *mx,
                  P7I_SPECIALS *sp, int *sexp, int *exp_out, float *mant_out) {
  int M = p->M;
  P7I_XF xf = p7i_xf_config(Lcfg, multihit);
  size_t W = (size_t)M + 2;
  float *buf = (float *)calloc(W * 8, sizeof(float));
  float *Mp = buf, *Ip = buf + W, *Dp = buf + 2 * W, *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;
  float *wA = buf + 6 * W, *wP = buf + 7 * W;
  float xN = 1.0f, xB = xf.tNB, xE = 0.0f, xJ = 0.0f, xC = 0.0f;
  int e_tot = 0;
  if (sp) { sp[0].xN = xN; sp[0].xB = xB; sp[0].xE = 0; sp[0].xJ = 0; sp[0].xC = 0; }
  if (sexp) sexp[0] = 0;
  if (mx) { memset(mx->M, 0, W * sizeof(float)); memset(mx->I, 0, W * sizeof(float)); memset(mx->D, 0, W * sizeof(float)); }
  for (int i = 1; i <= L; i++) {
    xE = p7i_fwd_row(p, dsq[i], Mp, Ip, Dp, xB, Mc, Ic, Dc, wA, wP);
    xN = xN * xf.tNN;
    xJ = xJ * xf.tJJ + xE * xf.tEJ;
    xC = xC * xf.tCC + xE * xf.tEC;
    xB = xJ * xf.tJB + xN * xf.tNB;
    if (xE > P7I_RESCALE_THRESH) {
      int e = ilogbf(xE);
      for (int k = 1; k <= M; k++) {
        Mc[k] = scalbnf(Mc[k], -e);
        Ic[k] = scalbnf(Ic[k], -e);
        Dc[k] = scalbnf(Dc[k], -e);
      }
      xN = scalbnf(xN, -e); xJ = scalbnf(xJ, -e); xC = scalbnf(xC, -e); xB = scalbnf(xB, -e); xE = scalbnf(xE, -e);
      e_tot += e;
    }
    if (sp) { sp[i].xN = xN; sp[i].xB = xB; sp[i].xE = xE; sp[i].xJ = xJ; sp[i].xC = xC; }
    if (sexp) sexp[i] = e_tot;
    if (mx) {
      memcpy(mx->M + (size_t)i * W, Mc, W * sizeof(float));
      memcpy(mx->I + (size_t)i * W, Ic, W * sizeof(float));
      memcpy(mx->D + (size_t)i * W, Dc, W * sizeof(float));
    }
    float *t;
    t = Mp; Mp = Mc; Mc = t;
    t = Ip; Ip = Ic; Ic = t;
    t = Dp; Dp = Dc; Dc = t;
  }
  float mant = xC * xf.tCT;
  free(buf);
  if (exp_out) *exp_out = e_tot;
  if (mant_out) *mant_out = mant;
  return (float)((double)e_tot * 0.69314718055994529 + log((double)mant));
}

/* Backward. Same storage conventions; sp[i]/sexp[i] for rows 0..L. Returns total (nats). */
float p7i_backward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, P7I_MX *mx,
                   P7I_SPECIALS *sp, int *sexp) {
  int M = p->M;
  P7I_XF xf = p7i_xf_config(Lcfg, multihit);
  size_t W = (size_t)M + 2;
  float *buf = (float *)calloc(W * 9, sizeof(float));
  float *Mn = buf, *In = buf + W, *Dn = buf + 2 * W, *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;
  float *we = buf + 6 * W, *wA = buf + 7 * W, *wP = buf + 8 * W;
  P7I_SPECIALS sn, sc;
  int e_tot = 0;
  sn.xC = xf.tCT; sn.xE = sn.xC * xf.tEC; sn.xJ = 0; sn.xB = 0; sn.xN = 0;
  /* row L: no next residue: we = 0, I(L,.) = 0; M_k -> E, M_k -> D_k+1 .. -> E, D_k -> E */
  p7i_bwd_cells(p, we, In, sn.xE, Mn, In, Dn, wA, wP); /* we, In are zero-filled (calloc) */
  if (sp) sp[L] = sn;
  if (sexp) sexp[L] = 0;
  if (mx) {
    memcpy(mx->M + (size_t)L * W, Mn, W * sizeof(float));
    memcpy(mx->I + (size_t)L * W, In, W * sizeof(float));
    memcpy(mx->D + (size_t)L * W, Dn, W * sizeof(float));
  }
  for (int i = L - 1; i >= 0; i--) {
    p7i_bwd_row(p, dsq[i + 1], Mn, In, &xf, &sn, &sc, Mc, Ic, Dc, we, wA, wP);
    float rowmax = 0.0f;
    for (int k = 1; k <= M; k++) if (Mc[k] > rowmax) rowmax = Mc[k];
    if (rowmax > P7I_RESCALE_THRESH) {
      int e = ilogbf(rowmax);
      for (int k = 1; k <= M; k++) {
        Mc[k] = scalbnf(Mc[k], -e);
        Ic[k] = scalbnf(Ic[k], -e);
        Dc[k] = scalbnf(Dc[k], -e);
      }
      sc.xN = scalbnf(sc.xN, -e); sc.xJ = scalbnf(sc.xJ, -e); sc.xC = scalbnf(sc.xC, -e);
      sc.xB = scalbnf(sc.xB, -e); sc.xE = scalbnf(sc.xE, -e);
      e_tot += e;
    }
    if (sp) sp[i] = sc;
    if (sexp) sexp[i] = e_tot;
    if (mx) {
      memcpy(mx->M + (size_t)i * W, Mc, W * sizeof(float));
      memcpy(mx->I + (size_t)i * W, Ic, W * sizeof(float));
      memcpy(mx->D + (size_t)i * W, Dc, W * sizeof(float));
    }
    sn = sc;
    float *t;
    t = Mn; Mn = Mc; Mc = t;
    t = In; In = Ic; Ic = t;
    t = Dn; Dn = Dc; Dc = t;
  }
  float tot = (float)((double)e_tot * 0.69314718055994529 + log((double)sn.xN));
  free(buf);
  return tot;
}

float orc_fwd_parser(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, int *exp_out,
                     float *mant_out) {
  return p7i_forward(p, dsq, L, Lcfg, multihit, NULL, NULL, NULL, exp_out, mant_out);
}
float orc_bwd_parser_total(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit) {
  return p7i_backward(p, dsq, L, Lcfg, multihit, NULL, NULL, NULL);
}

/* ---------------- textbook log-space cross-checks (double) ---------------- */
static double lse2(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

static double generic_dp(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, int viterbi) {
  int M = p->M;
  double nj = multihit ? 1.0 : 0.0;
  double pmove = (2.0 + nj) / ((double)Lcfg + 2.0 + nj), ploop = 1.0 - pmove;
  double tloop = log(ploop), tmove = log(pmove);
  double tEJ = multihit ? log(0.5) : -INFINITY, tEC = multihit ? log(0.5) : 0.0;
  size_t W = (size_t)M + 1;
  double *buf = (double *)malloc(sizeof(double) * W * 6);
  double *Mp = buf, *Ip = buf + W, *Dp = buf + 2 * W, *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;
  for (size_t k = 0; k < W * 6; k++) buf[k] = -INFINITY;
  double xN = 0.0, xB = tmove, xJ = -INFINITY, xC = -INFINITY, xE;
#define COMB(a, b) (viterbi ? ((a) > (b) ? (a) : (b)) : lse2((a), (b)))
  for (int i = 1; i <= L; i++) {
    const float *rs = p->msc + (size_t)dsq[i] * (M + 1);
    xE = -INFINITY;
    Mc[0] = Ic[0] = Dc[0] = -INFINITY;
    for (int k = 1; k <= M; k++) {
      const float *t = p->tsc + (k - 1) * 8;
      double sv = xB + t[ORC_BM];
      sv = COMB(sv, Mp[k - 1] + t[ORC_MM]);
      sv = COMB(sv, Ip[k - 1] + t[ORC_IM]);
      sv = COMB(sv, Dp[k - 1] + t[ORC_DM]);
      Mc[k] = sv + rs[k];
      Dc[k] = COMB(Mc[k - 1] + t[ORC_MD], Dc[k - 1] + t[ORC_DD]);
      const float *tk = p->tsc + k * 8;
      Ic[k] = (k < M) ? COMB(Mp[k] + tk[ORC_MI], Ip[k] + tk[ORC_II]) : -INFINITY;
      xE = COMB(xE, Mc[k]);
      if (!viterbi || 1) xE = COMB(xE, Dc[k]);
    }
    xN = xN + tloop;
    xJ = COMB(xJ + tloop, xE + tEJ);
    xC = COMB(xC + tloop, xE + tEC);
    xB = COMB(xJ + tmove, xN + tmove);
    double *t;
    t = Mp; Mp = Mc; Mc = t;
    t = Ip; Ip = Ic; Ic = t;
    t = Dp; Dp = Dc; Dc = t;
  }
  free(buf);
  return xC + tmove;
}
double orc_generic_forward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit) {
  return generic_dp(p, dsq, L, Lcfg, multihit, 0);
}
double orc_generic_viterbi(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit) {
  return generic_dp(p, dsq, L, Lcfg, multihit, 1);
}
