/*
 * p7_domain.c -- ORACLE (test infrastructure): posterior domain definition,
 * envelope rescoring (unihit Forward/Backward), posterior decoding, optimal
 * accuracy alignment + traceback, null2 by expectation, alignment display.
 * Restates SURVEY.md Appendix A.7, A.8, A.10 (HMMER3 domain definition; the
 * reference consumes its output at /root/reference/lib/aln.py:511-561).
 *
 * Multi-domain regions (rt3) are resolved the way HMMER does (A.7): 200 stochastic tracebacks of the
 * multihit Forward matrix of the region, single-linkage clustering of the sampled domain segments,
 * clusters seen in >= 25 % of the traces become envelopes whose ends are the outermost end points used
 * by >= 2 % of the traces.  Two documented deviations: the random stream is this file's own
 * (xorshift64*, one stream per trace index, so the result does not depend on evaluation order; HMMER's
 * is its Mersenne Twister reseeded per target and cannot be reproduced without its source), and null2 of
 * such envelopes is computed by expectation like every other envelope instead of from the traces.
 */
#include "p7_domain.h"
#include "p7_internal.h"
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#pragma STDC FP_CONTRACT OFF

static P7I_MX mx_alloc(int L, int M) {
  P7I_MX m;
  size_t n = (size_t)(L + 1) * (M + 2);
  m.M = (float *)calloc(n, sizeof(float));
  m.I = (float *)calloc(n, sizeof(float));
  m.D = (float *)calloc(n, sizeof(float));
  return m;
}
static void mx_free(P7I_MX *m) {
  free(m->M);
  free(m->I);
  free(m->D);
}

/* btot/etot/mocc from multihit Forward/Backward parsers (A.7); arrays [0..L] */
int orc_domain_decoding(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *btot, float *etot,
                        float *mocc) {
  P7I_SPECIALS *sf = (P7I_SPECIALS *)malloc(sizeof(P7I_SPECIALS) * (L + 1));
  P7I_SPECIALS *sb = (P7I_SPECIALS *)malloc(sizeof(P7I_SPECIALS) * (L + 1));
  int *ef = (int *)malloc(sizeof(int) * (L + 1)), *eb = (int *)malloc(sizeof(int) * (L + 1));
  int Ze;
  float Zm;
  P7I_XF xf = p7i_xf_config(L, 1);
  p7i_forward(p, dsq, L, L, 1, NULL, sf, ef, &Ze, &Zm);
  p7i_backward(p, dsq, L, L, 1, NULL, sb, eb);
  btot[0] = etot[0] = mocc[0] = 0.0f;
  for (int i = 1; i <= L; i++) {
    float pb = scalbnf(sf[i - 1].xB * sb[i - 1].xB, ef[i - 1] + eb[i - 1] - Ze) / Zm;
    btot[i] = btot[i - 1] + pb;
    float pe = scalbnf(sf[i].xE * sb[i].xE, ef[i] + eb[i] - Ze) / Zm;
    etot[i] = etot[i - 1] + pe;
    int ex = ef[i - 1] + eb[i] - Ze;
    float n = scalbnf((sf[i - 1].xN * sb[i].xN) * xf.tNN, ex) / Zm;
    float j = scalbnf((sf[i - 1].xJ * sb[i].xJ) * xf.tJJ, ex) / Zm;
    float c = scalbnf((sf[i - 1].xC * sb[i].xC) * xf.tCC, ex) / Zm;
    float njcp = (n + j) + c;
    mocc[i] = 1.0f - njcp;
  }
  free(sf);
  free(sb);
  free(ef);
  free(eb);
  return 0;
}

static int is_multidomain_region(const float *btot, const float *etot, int i, int j, float rt3) {
  float max = -1.0f;
  for (int z = i; z <= j; z++) {
    float a = etot[z] - etot[i - 1], b = btot[j] - btot[z - 1];
    float expected_n = a < b ? a : b;
    if (expected_n > max) max = expected_n;
  }
  return max >= rt3;
}

/* ---- envelope rescoring ---- */
typedef struct {
  int n, cap;
  char *st;
  int *k, *i;
} TRACE;
static void tr_push(TRACE *t, char st, int k, int i) {
  if (t->n == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 256;
    t->st = (char *)realloc(t->st, t->cap);
    t->k = (int *)realloc(t->k, sizeof(int) * t->cap);
    t->i = (int *)realloc(t->i, sizeof(int) * t->cap);
  }
  t->st[t->n] = st;
  t->k[t->n] = k;
  t->i[t->n] = i;
  t->n++;
}

/* Rescore dsq[ienv..jenv] (1-based, inside the full target dsq[1..L]) in unihit mode with the
 * length model of the full target. Fills dom; n2sc[ienv..jenv] set (null2 by expectation).
 * model/target receive HMMER-style alignment lines (NUL-terminated); the posterior-probability line (HMMER's "PP"
 * annotation: p7_alidisplay_EncodePostProb, '.' for deletes) follows target's NUL, so a domain takes 2*(ali_len+1)
 * bytes of each pool. Returns 0, or 1 if no M state. */
/* HMMER's p7_alidisplay_EncodePostProb */
static char p7d_pp_char(float p) { return (p + 0.05 >= 1.0) ? '*' : (char)((int)((p + 0.05) * 10.0) + '0'); }

int p7d_rescore_envelope(const ORC_PROFILE *p, const uint8_t *dsq, int L, int ienv, int jenv, int do_null2,
                         float *n2sc, P7D_DOMAIN *dom, char *model, char *target, int alicap) {
  int M = p->M, Ld = jenv - ienv + 1;
  size_t W = (size_t)M + 2;
  const uint8_t *sub = dsq + ienv - 1; /* sub[1..Ld] */
  P7I_MX fx = mx_alloc(Ld, M), bx = mx_alloc(Ld, M);
  P7I_SPECIALS *sf = (P7I_SPECIALS *)malloc(sizeof(P7I_SPECIALS) * (Ld + 1));
  P7I_SPECIALS *sb = (P7I_SPECIALS *)malloc(sizeof(P7I_SPECIALS) * (Ld + 1));
  int *ef = (int *)malloc(sizeof(int) * (Ld + 1)), *eb = (int *)malloc(sizeof(int) * (Ld + 1));
  int Ze;
  float Zm;
  P7I_XF xf = p7i_xf_config(L, 0);
  const float *tf = p->tfv;
  float envsc = p7i_forward(p, sub, Ld, L, 0, &fx, sf, ef, &Ze, &Zm);
  p7i_backward(p, sub, Ld, L, 0, &bx, sb, eb);

  /* posterior decoding: overwrite bx with pp (D cells unused = 0); specials in ppN/ppJ/ppC */
  float *ppN = (float *)calloc((size_t)Ld + 1, sizeof(float));
  float *ppJ = (float *)calloc((size_t)Ld + 1, sizeof(float));
  float *ppC = (float *)calloc((size_t)Ld + 1, sizeof(float));
  for (int i = 1; i <= Ld; i++) {
    int ex = ef[i] + eb[i] - Ze;
    float *fM = fx.M + i * W, *fI = fx.I + i * W, *bM = bx.M + i * W, *bI = bx.I + i * W, *bD = bx.D + i * W;
    for (int k = 1; k <= M; k++) {
      bM[k] = scalbnf(fM[k] * bM[k], ex) / Zm;
      bI[k] = scalbnf(fI[k] * bI[k], ex) / Zm;
      bD[k] = 0.0f;
    }
    int ex2 = ef[i - 1] + eb[i] - Ze;
    ppN[i] = scalbnf((sf[i - 1].xN * sb[i].xN) * xf.tNN, ex2) / Zm;
    ppJ[i] = scalbnf((sf[i - 1].xJ * sb[i].xJ) * xf.tJJ, ex2) / Zm;
    ppC[i] = scalbnf((sf[i - 1].xC * sb[i].xC) * xf.tCC, ex2) / Zm;
  }

  /* optimal accuracy DP: overwrite fx with OA values */
  float *oE = (float *)malloc(sizeof(float) * (Ld + 1)), *oN = (float *)malloc(sizeof(float) * (Ld + 1));
  float *oJ = (float *)malloc(sizeof(float) * (Ld + 1)), *oB = (float *)malloc(sizeof(float) * (Ld + 1));
  float *oC = (float *)malloc(sizeof(float) * (Ld + 1));
  for (size_t z = 0; z < W; z++) fx.M[z] = fx.I[z] = fx.D[z] = -INFINITY; /* row 0: no residue yet */
  oN[0] = 0.0f;
  oB[0] = 0.0f;
  oE[0] = oJ[0] = oC[0] = -INFINITY;
#define OK_(k, t) (tf[(k) * 8 + (t)] > 0.0f)
  for (int i = 1; i <= Ld; i++) {
    float *Mp = fx.M + (i - 1) * W, *Ip = fx.I + (i - 1) * W, *Dp = fx.D + (i - 1) * W;
    float *Mc = fx.M + i * W, *Ic = fx.I + i * W, *Dc = fx.D + i * W;
    const float *pM = bx.M + i * W, *pI = bx.I + i * W;
    float xE = 0.0f; /* impl_sse/optacc: xEv starts at zero */
    Mc[0] = Ic[0] = Dc[0] = 0.0f;
    for (int k = 1; k <= M; k++) {
      float sv = OK_(k - 1, ORC_BM) ? oB[i - 1] : 0.0f;
      float v;
      v = OK_(k - 1, ORC_MM) ? Mp[k - 1] : 0.0f; if (v > sv) sv = v;
      v = OK_(k - 1, ORC_IM) ? Ip[k - 1] : 0.0f; if (v > sv) sv = v;
      v = OK_(k - 1, ORC_DM) ? Dp[k - 1] : 0.0f; if (v > sv) sv = v;
      Mc[k] = sv + pM[k];
      if (Mc[k] > xE) xE = Mc[k];
      float a = OK_(k - 1, ORC_MD) ? Mc[k - 1] : 0.0f;
      float b = OK_(k - 1, ORC_DD) ? Dc[k - 1] : 0.0f;
      Dc[k] = a > b ? a : b;
      a = OK_(k, ORC_MI) ? Mp[k] : 0.0f;
      b = OK_(k, ORC_II) ? Ip[k] : 0.0f;
      Ic[k] = (a > b ? a : b) + pI[k];
    }
    for (int k = 1; k <= M; k++) if (Dc[k] > xE) xE = Dc[k];
    oE[i] = xE;
    float t1, t2;
    t1 = (xf.tJJ == 0.0f) ? 0.0f : oJ[i - 1] + ppJ[i];
    t2 = (xf.tEJ == 0.0f) ? 0.0f : oE[i];
    /* J unreachable in unihit mode: keep it at -inf so a zero N never ties with it */
    oJ[i] = (xf.tJJ == 0.0f && xf.tEJ == 0.0f) ? -INFINITY : (t1 > t2 ? t1 : t2);
    t1 = (xf.tCC == 0.0f) ? 0.0f : oC[i - 1] + ppC[i];
    t2 = (xf.tEC == 0.0f) ? 0.0f : oE[i];
    oC[i] = t1 > t2 ? t1 : t2;
    oN[i] = (xf.tNN == 0.0f) ? 0.0f : oN[i - 1] + ppN[i];
    t1 = (xf.tNB == 0.0f) ? 0.0f : oN[i];
    t2 = (xf.tJB == 0.0f) ? 0.0f : oJ[i];
    oB[i] = t1 > t2 ? t1 : t2;
  }
  float oasc = oC[Ld];

  /* OA traceback (built backwards) */
  TRACE tr = {0, 0, NULL, NULL, NULL};
  {
    int i = Ld, k = 0;
    char sprv = 'C', scur;
    tr_push(&tr, 'T', 0, 0);
    tr_push(&tr, 'C', 0, 0);
    int guard = 0, fail = 0;
    while (sprv != 'S' && !fail) {
      switch (sprv) {
        case 'M': {
          float path[4];
          path[0] = OK_(k - 1, ORC_MM) ? fx.M[(i - 1) * W + k - 1] : -INFINITY;
          path[1] = OK_(k - 1, ORC_IM) ? fx.I[(i - 1) * W + k - 1] : -INFINITY;
          path[2] = OK_(k - 1, ORC_DM) ? fx.D[(i - 1) * W + k - 1] : -INFINITY;
          path[3] = OK_(k - 1, ORC_BM) ? oB[i - 1] : -INFINITY;
          int best = 0;
          for (int z = 1; z < 4; z++) if (path[z] > path[best]) best = z;
          scur = "MIDB"[best];
          k--; i--;
        } break;
        case 'D': {
          float a = OK_(k - 1, ORC_MD) ? fx.M[i * W + k - 1] : -INFINITY;
          float b = OK_(k - 1, ORC_DD) ? fx.D[i * W + k - 1] : -INFINITY;
          scur = (a >= b) ? 'M' : 'D';
          k--;
        } break;
        case 'I': {
          float a = OK_(k, ORC_MI) ? fx.M[(i - 1) * W + k] : -INFINITY;
          float b = OK_(k, ORC_II) ? fx.I[(i - 1) * W + k] : -INFINITY;
          scur = (a >= b) ? 'M' : 'I';
          i--;
        } break;
        case 'N': scur = (i == 0) ? 'S' : 'N'; break;
        case 'C': {
          float a = (xf.tCC == 0.0f) ? -INFINITY : oC[i - 1] + ppC[i];
          float b = (xf.tEC == 0.0f) ? -INFINITY : oE[i];
          if (i == 0) { fail = 1; scur = 'S'; break; }
          scur = (a > b) ? 'C' : 'E';
        } break;
        case 'J': {
          float a = (xf.tJJ == 0.0f) ? -INFINITY : oJ[i - 1] + ppJ[i];
          float b = (xf.tEJ == 0.0f) ? -INFINITY : oE[i];
          scur = (a > b) ? 'J' : 'E';
        } break;
        case 'E': {
          float max = -INFINITY;
          int kmax = 0;
          char smax = 'M';
          for (int kk = 1; kk <= M; kk++) {
            if (fx.M[i * W + kk] > max) { max = fx.M[i * W + kk]; kmax = kk; smax = 'M'; }
            if (fx.D[i * W + kk] > max) { max = fx.D[i * W + kk]; kmax = kk; smax = 'D'; }
          }
          k = kmax;
          scur = smax;
        } break;
        case 'B': {
          float a = (xf.tNB == 0.0f) ? -INFINITY : oN[i];
          float b = (xf.tJB == 0.0f) ? -INFINITY : oJ[i];
          scur = (a > b) ? 'N' : 'J';
        } break;
        default: fail = 1; scur = 'S';
      }
      if (scur == 'S') { tr_push(&tr, 'S', 0, 0); break; }
      /* emitted-residue index convention: M,I carry i of the emitted residue; N/J/C emit on loop */
      if (scur == 'M' || scur == 'I' || scur == 'D') tr_push(&tr, scur, k, (scur == 'D') ? 0 : i);
      else if ((scur == 'N' || scur == 'J' || scur == 'C') && scur == sprv) { tr_push(&tr, scur, 0, i); i--; }
      else tr_push(&tr, scur, 0, 0);
      sprv = scur;
      if (++guard > 4 * (Ld + M) + 16) fail = 1;
    }
    if (fail) tr.n = 0;
  }
  /* reverse */
  for (int a = 0, b = tr.n - 1; a < b; a++, b--) {
    char c = tr.st[a]; tr.st[a] = tr.st[b]; tr.st[b] = c;
    int t = tr.k[a]; tr.k[a] = tr.k[b]; tr.k[b] = t;
    t = tr.i[a]; tr.i[a] = tr.i[b]; tr.i[b] = t;
  }
  /* Each pushed state carries its own coordinates: the predecessor selection decrements (k,i) to
   * the predecessor's cell before it is appended, as HMMER's p7_OATrace does. */

  /* null2 by expectation over the envelope */
  float domcorr = 0.0f;
  if (do_null2) {
    int K = p->K, Kp = p->Kp, B = (M + ORC_NL - 1) / ORC_NL, mem[20];
    float *eM = (float *)calloc(W, sizeof(float)), *eI = (float *)calloc(W, sizeof(float));
    float sN = 0.0f, sJ = 0.0f, sC = 0.0f;
    for (int i = 1; i <= Ld; i++) {
      const float *pM = bx.M + i * W, *pI = bx.I + i * W;
      for (int k = 1; k <= M; k++) {
        eM[k] = eM[k] + pM[k];
        eI[k] = eI[k] + pI[k];
      }
      sN = sN + ppN[i];
      sJ = sJ + ppJ[i];
      sC = sC + ppC[i];
    }
    float norm = 1.0f / (float)Ld;
    for (int k = 1; k <= M; k++) {
      eM[k] = eM[k] * norm;
      eI[k] = eI[k] * norm;
    }
    float xfactor = (sN * norm + sC * norm) + sJ * norm;
    float null2[32];
    for (int x = 0; x < K; x++) {
      const float *rf = p->rfv + (size_t)x * (M + 1);
      float lane[ORC_NL];
      for (int l = 0; l < ORC_NL; l++) {
        int first = l * B + 1, last = (l + 1) * B;
        if (last > M) last = M;
        float s = 0.0f;
        for (int k = first; k <= last; k++) {
          s = s + eM[k] * rf[k];
          s = s + eI[k];
        }
        lane[l] = s;
      }
      null2[x] = p7i_lane_sum(lane) + xfactor;
    }
    for (int x = K + 1; x <= Kp - 3; x++) {
      p7i_degen_members(p->abc, x, mem);
      float s = 0.0f;
      int n = 0;
      for (int y = 0; y < K; y++)
        if (mem[y]) { s = s + null2[y]; n++; }
      null2[x] = s / (float)n;
    }
    null2[K] = null2[Kp - 2] = null2[Kp - 1] = 1.0f;
    for (int pos = ienv; pos <= jenv; pos++) {
      n2sc[pos] = (float)log((double)null2[dsq[pos]]);
      domcorr += n2sc[pos];
    }
    free(eM);
    free(eI);
  }

  /* alignment display from the trace: first..last M state */
  int z1 = -1, z2 = -1, rc = 0;
  for (int z = 0; z < tr.n; z++) if (tr.st[z] == 'M') { z1 = z; break; }
  for (int z = tr.n - 1; z >= 0; z--) if (tr.st[z] == 'M') { z2 = z; break; }
  dom->ienv = ienv;
  dom->jenv = jenv;
  dom->envsc = envsc;
  dom->oasc = oasc;
  dom->domcorrection = domcorr;
  if (z1 < 0) {
    rc = 1;
    dom->ali_len = 0;
    model[0] = target[0] = 0;
    if (alicap > 1) model[1] = target[1] = 0;
  } else {
    int n = 0;
    dom->hmmfrom = tr.k[z1];
    dom->hmmto = tr.k[z2];
    dom->iali = tr.i[z1] + ienv - 1;
    dom->jali = tr.i[z2] + ienv - 1;
    char *pp = (char *)malloc((size_t)alicap);
    for (int z = z1; z <= z2 && 2 * (n + 1) < alicap - 1; z++) {
      char st = tr.st[z];
      if (st == 'M') {
        model[n] = p->hmm->cons[tr.k[z]];
        target[n] = (char)toupper((unsigned char)orc_sym(p->abc, sub[tr.i[z]]));
        pp[n] = p7d_pp_char(bx.M[(size_t)tr.i[z] * W + tr.k[z]]);
      } else if (st == 'I') {
        model[n] = '.';
        target[n] = (char)tolower((unsigned char)orc_sym(p->abc, sub[tr.i[z]]));
        pp[n] = p7d_pp_char(bx.I[(size_t)tr.i[z] * W + tr.k[z]]);
      } else if (st == 'D') {
        model[n] = p->hmm->cons[tr.k[z]];
        target[n] = '-';
        pp[n] = '.';
      } else
        continue;
      n++;
    }
    model[n] = target[n] = 0;
    memcpy(target + n + 1, pp, (size_t)n);
    target[2 * n + 1] = 0;
    memset(model + n + 1, 0, (size_t)n + 1);
    free(pp);
    dom->ali_len = n;
  }
  free(tr.st); free(tr.k); free(tr.i);
  free(oE); free(oN); free(oJ); free(oB); free(oC);
  free(ppN); free(ppJ); free(ppC);
  free(sf); free(sb); free(ef); free(eb);
  mx_free(&fx);
  mx_free(&bx);
  return rc;
}

/* ---- multi-domain regions: stochastic traceback ensemble + clustering (A.7) ---- */
#define P7D_NSAMPLES 200
#define P7D_MAXSEG 16 /* domain segments kept per sampled trace (further ones are dropped) */
typedef struct { int t, i, j, k, m; } P7D_SEG;

static inline float p7d_rng_uniform(uint64_t *x) { /* xorshift64*, top 24 bits -> [0,1) */
  *x ^= *x >> 12;
  *x ^= *x << 25;
  *x ^= *x >> 27;
  return (float)((*x * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
}
static inline uint64_t p7d_rng_seed(int t) { return 42ull * 0x9E3779B97F4A7C15ull + (uint64_t)(t + 1) * 0xBF58476D1CE4E5B9ull; }

/* index drawn with probability p[i] / sum(p): serial sums in index order */
static int p7d_choose(const float *pr, int n, float u) {
  float tot = 0.0f;
  for (int i = 0; i < n; i++) tot = tot + pr[i];
  const float r = u * tot;
  float cum = 0.0f;
  int last = 0;
  for (int i = 0; i < n; i++) {
    if (pr[i] > 0.0f) last = i;
    cum = cum + pr[i];
    if (r < cum) return i;
  }
  return last;
}

/* One stochastic traceback of the multihit Forward matrix of a region (rows 0..Lr). Appends the
 * trace's domain segments (first/last match state: residue i..j, node k..m). */
static int p7d_sample_trace(const ORC_PROFILE *p, const P7I_MX *fx, const P7I_SPECIALS *sf, const int *ef, int Lr,
                            const P7I_XF *xf, int t, P7D_SEG *out) {
  const int M = p->M;
  const size_t W = (size_t)M + 2;
  const float *tf = p->tfv;
  uint64_t rng = p7d_rng_seed(t);
  int nseg = 0, i = Lr, k = 0;
  int seg_j = 0, seg_m = 0, seg_i = 0, seg_k = 0, have_m = 0;
  char st = 'C';
  int guard = 0;
  while (st != 'S' && guard++ < 8 * (Lr + M) + 64) {
    float pr[4];
    switch (st) {
      case 'C':
        if (i == 0) { st = 'S'; break; } /* cannot happen with positive probability */
        pr[0] = scalbnf(sf[i - 1].xC * xf->tCC, ef[i - 1] - ef[i]);
        pr[1] = sf[i].xE * xf->tEC;
        if (p7d_choose(pr, 2, p7d_rng_uniform(&rng)) == 0) i--;
        else st = 'E';
        break;
      case 'J':
        if (i == 0) { st = 'S'; break; }
        pr[0] = scalbnf(sf[i - 1].xJ * xf->tJJ, ef[i - 1] - ef[i]);
        pr[1] = sf[i].xE * xf->tEJ;
        if (p7d_choose(pr, 2, p7d_rng_uniform(&rng)) == 0) i--;
        else st = 'E';
        break;
      case 'E': { /* all M_k, D_k of row i, k ascending, M before D */
        const float *Mi = fx->M + (size_t)i * W, *Di = fx->D + (size_t)i * W;
        float tot = 0.0f;
        for (int kk = 1; kk <= M; kk++) { tot = tot + Mi[kk]; tot = tot + Di[kk]; }
        const float r = p7d_rng_uniform(&rng) * tot;
        float cum = 0.0f;
        int ck = M, cs = 'M', found = 0;
        for (int kk = 1; kk <= M && !found; kk++) {
          if (Mi[kk] > 0.0f) { ck = kk; cs = 'M'; }
          cum = cum + Mi[kk];
          if (r < cum) { ck = kk; cs = 'M'; found = 1; break; }
          if (Di[kk] > 0.0f) { ck = kk; cs = 'D'; }
          cum = cum + Di[kk];
          if (r < cum) { ck = kk; cs = 'D'; found = 1; break; }
        }
        k = ck;
        st = (char)cs;
        have_m = 0;
      } break;
      case 'M': {
        if (!have_m) { seg_j = i; seg_m = k; have_m = 1; }
        seg_i = i; seg_k = k;
        const float *Mp = fx->M + (size_t)(i - 1) * W, *Ip = fx->I + (size_t)(i - 1) * W, *Dp = fx->D + (size_t)(i - 1) * W;
        pr[0] = sf[i - 1].xB * tf[(k - 1) * 8 + ORC_BM];
        pr[1] = k >= 2 ? Mp[k - 1] * tf[(k - 1) * 8 + ORC_MM] : 0.0f;
        pr[2] = k >= 2 ? Ip[k - 1] * tf[(k - 1) * 8 + ORC_IM] : 0.0f;
        pr[3] = k >= 2 ? Dp[k - 1] * tf[(k - 1) * 8 + ORC_DM] : 0.0f;
        const int c = p7d_choose(pr, 4, p7d_rng_uniform(&rng));
        i--;
        if (c == 0) st = 'B';
        else { k--; st = c == 1 ? 'M' : c == 2 ? 'I' : 'D'; }
      } break;
      case 'D': {
        const float *Mi = fx->M + (size_t)i * W, *Di = fx->D + (size_t)i * W;
        pr[0] = k >= 2 ? Mi[k - 1] * tf[(k - 1) * 8 + ORC_MD] : 0.0f;
        pr[1] = k >= 2 ? Di[k - 1] * tf[(k - 1) * 8 + ORC_DD] : 0.0f;
        const int c = p7d_choose(pr, 2, p7d_rng_uniform(&rng));
        k--;
        st = c == 0 ? 'M' : 'D';
      } break;
      case 'I': {
        const float *Mp = fx->M + (size_t)(i - 1) * W, *Ip = fx->I + (size_t)(i - 1) * W;
        pr[0] = Mp[k] * tf[k * 8 + ORC_MI];
        pr[1] = Ip[k] * tf[k * 8 + ORC_II];
        const int c = p7d_choose(pr, 2, p7d_rng_uniform(&rng));
        i--;
        st = c == 0 ? 'M' : 'I';
      } break;
      case 'B': {
        if (have_m && nseg < P7D_MAXSEG) { out[nseg].t = t; out[nseg].i = seg_i; out[nseg].j = seg_j; out[nseg].k = seg_k; out[nseg].m = seg_m; nseg++; }
        have_m = 0;
        pr[0] = sf[i].xN * xf->tNB;
        pr[1] = sf[i].xJ * xf->tJB;
        st = p7d_choose(pr, 2, p7d_rng_uniform(&rng)) == 0 ? 'N' : 'J';
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

static int p7d_linked(const P7D_SEG *a, const P7D_SEG *b) { /* min_overlap 0.8 of the smaller, max_diagdiff 4 */
  int nov = (a->j < b->j ? a->j : b->j) - (a->i > b->i ? a->i : b->i) + 1;
  int la = a->j - a->i + 1, lb = b->j - b->i + 1;
  int n = la < lb ? la : lb;
  if ((float)nov / (float)n < 0.8f) return 0;
  nov = (a->m < b->m ? a->m : b->m) - (a->k > b->k ? a->k : b->k) + 1;
  la = a->m - a->k + 1; lb = b->m - b->k + 1;
  n = la < lb ? la : lb;
  if ((float)nov / (float)n < 0.8f) return 0;
  if (abs((a->k - a->i) - (b->k - b->i)) > 4) return 0;
  if (abs((a->m - a->j) - (b->m - b->j)) > 4) return 0;
  return 1;
}

/* Envelopes (target coordinates) of a multi-domain region ireg..jreg; returns their number (<= maxenv),
 * ordered by start, then end. */
int p7d_resolve_multidomain(const ORC_PROFILE *p, const uint8_t *dsq, int L, int ireg, int jreg, int *env_i, int *env_j,
                            int maxenv) {
  const int M = p->M, Lr = jreg - ireg + 1;
  const uint8_t *sub = dsq + ireg - 1;
  P7I_MX fx = mx_alloc(Lr, M);
  P7I_SPECIALS *sf = (P7I_SPECIALS *)malloc(sizeof(P7I_SPECIALS) * (Lr + 1));
  int *ef = (int *)malloc(sizeof(int) * (Lr + 1));
  P7I_XF xf = p7i_xf_config(L, 1);
  p7i_forward(p, sub, Lr, L, 1, &fx, sf, ef, NULL, NULL);
  P7D_SEG *seg = (P7D_SEG *)malloc(sizeof(P7D_SEG) * P7D_NSAMPLES * P7D_MAXSEG);
  int nseg = 0;
  for (int t = 0; t < P7D_NSAMPLES; t++) nseg += p7d_sample_trace(p, &fx, sf, ef, Lr, &xf, t, seg + nseg);
  /* single linkage: connected components by label propagation (label = smallest member index) */
  int *lab = (int *)malloc(sizeof(int) * (nseg + 1));
  for (int a = 0; a < nseg; a++) lab[a] = a;
  for (int changed = 1; changed;) {
    changed = 0;
    for (int a = 0; a < nseg; a++)
      for (int b = 0; b < nseg; b++)
        if (lab[b] < lab[a] && p7d_linked(&seg[a], &seg[b])) { lab[a] = lab[b]; changed = 1; }
  }
  int nenv = 0;
  int *cnt_i = (int *)calloc((size_t)Lr + 2, sizeof(int)), *cnt_j = (int *)calloc((size_t)Lr + 2, sizeof(int));
  for (int c = 0; c < nseg; c++) {
    if (lab[c] != c) continue; /* c is the representative (smallest index) of its cluster */
    int ninc = 0, lastt = -1, imin = Lr + 1, imax = 0, jmin = Lr + 1, jmax = 0;
    for (int a = c; a < nseg; a++) {
      if (lab[a] != c) continue;
      if (seg[a].t != lastt) { ninc++; lastt = seg[a].t; } /* segments are stored trace by trace */
      cnt_i[seg[a].i]++;
      cnt_j[seg[a].j]++;
      if (seg[a].i < imin) imin = seg[a].i;
      if (seg[a].i > imax) imax = seg[a].i;
      if (seg[a].j < jmin) jmin = seg[a].j;
      if (seg[a].j > jmax) jmax = seg[a].j;
    }
    if ((float)ninc / (float)P7D_NSAMPLES >= 0.25f) {
      int bi, bj, arg = imin;
      for (bi = imin; bi <= imax; bi++) {
        if (cnt_i[bi] > cnt_i[arg]) arg = bi;
        if ((float)cnt_i[bi] / (float)ninc >= 0.02f) break;
      }
      if (bi > imax) bi = arg;
      arg = jmax;
      for (bj = jmax; bj >= jmin; bj--) {
        if (cnt_j[bj] > cnt_j[arg]) arg = bj;
        if ((float)cnt_j[bj] / (float)ninc >= 0.02f) break;
      }
      if (bj < jmin) bj = arg;
      if (nenv < maxenv && bj >= bi) { env_i[nenv] = bi + ireg - 1; env_j[nenv] = bj + ireg - 1; nenv++; }
    }
    for (int a = c; a < nseg; a++)
      if (lab[a] == c) { cnt_i[seg[a].i] = 0; cnt_j[seg[a].j] = 0; }
  }
  /* order by start, then end (insertion sort; nenv is tiny) */
  for (int a = 1; a < nenv; a++)
    for (int b = a; b > 0 && (env_i[b] < env_i[b - 1] || (env_i[b] == env_i[b - 1] && env_j[b] < env_j[b - 1])); b--) {
      int ti = env_i[b], tj = env_j[b];
      env_i[b] = env_i[b - 1]; env_j[b] = env_j[b - 1];
      env_i[b - 1] = ti; env_j[b - 1] = tj;
    }
  free(cnt_i); free(cnt_j); free(lab); free(seg); free(sf); free(ef);
  mx_free(&fx);
  return nenv;
}

/* Domain definition over a full target (A.7). doms[] capacity maxdom; alignment strings are written
 * consecutively (NUL-separated) into model/target pools. Returns ndom; *nregions, *nmulti set. */
int p7d_define_domains(const ORC_PROFILE *p, const uint8_t *dsq, int L, int do_null2, float *n2sc,
                       P7D_DOMAIN *doms, int maxdom, char *model, char *target, int poolcap, int *nregions,
                       int *nmulti) {
  const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
  float *btot = (float *)malloc(sizeof(float) * (L + 1));
  float *etot = (float *)malloc(sizeof(float) * (L + 1));
  float *mocc = (float *)malloc(sizeof(float) * (L + 1));
  int ndom = 0, used = 0;
  *nregions = 0;
  *nmulti = 0;
  orc_domain_decoding(p, dsq, L, btot, etot, mocc);
  for (int i = 0; i <= L; i++) n2sc[i] = 0.0f;
  int i = -1, triggered = 0;
  for (int j = 1; j <= L; j++) {
    if (!triggered) {
      if (mocc[j] - (btot[j] - btot[j - 1]) < rt2) i = j;
      else if (i == -1) i = j;
      if (mocc[j] >= rt1) triggered = 1;
    } else if (mocc[j] - (etot[j] - etot[j - 1]) < rt2) {
      (*nregions)++;
      int multi = is_multidomain_region(btot, etot, i, j, rt3);
      if (multi) (*nmulti)++;
      int env_i[P7D_MAXENV], env_j[P7D_MAXENV], nenv = 1;
      env_i[0] = i; env_j[0] = j;
      if (multi) nenv = p7d_resolve_multidomain(p, dsq, L, i, j, env_i, env_j, P7D_MAXENV);
      for (int e = 0; e < nenv; e++)
        if (ndom < maxdom && poolcap - used > 4 * (env_j[e] - env_i[e] + 1) + 4 * p->M + 16) {
          P7D_DOMAIN *d = &doms[ndom];
          d->multidomain_region = multi;
          d->ali_off = used;
          int rc = p7d_rescore_envelope(p, dsq, L, env_i[e], env_j[e], do_null2, n2sc, d, model + used, target + used, poolcap - used);
          if (rc == 0) {
            used += 2 * (d->ali_len + 1);
            ndom++;
          }
        }
      i = -1;
      triggered = 0;
    }
  }
  free(btot);
  free(etot);
  free(mocc);
  return ndom;
}
