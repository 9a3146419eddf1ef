/*
 * p7_pipeline.c -- ORACLE (test infrastructure): the hmmsearch acceleration
 * pipeline per target (MSV -> bias -> Viterbi -> Forward -> domain definition ->
 * scores); reference call site
 * /root/reference/lib/aln.py:664-667.  Z/domZ thresholds and the report order
 * (A.9, A.10) are applied by the caller after all targets are gathered.
 */
#include "p7oracle.h"
#include "p7_internal.h"
#include "p7_domain.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#pragma STDC FP_CONTRACT OFF
#define LN2 0.69314718055994529
#define MAXDOM_PER_SEQ 64

typedef struct {
  ORC_HIT *hits;
  int n, cap;
  char *model, *target;
  int used, poolcap;
} HITBUF;

static int one_target(const ORC_PROFILE *p, const uint8_t *seq, int L, int tidx, const ORC_OPTS *o,
                      ORC_TRACE *tr, ORC_HIT *out, int maxout, char *model, char *target, int poolcap,
                      int *pool_used) {
  const float *ev = p->hmm->ev;
  ORC_TRACE t;
  memset(&t, 0, sizeof(t));
  t.msv_xJ = 0;
  t.vit_xC = -32768;
  int nh = 0;
  *pool_used = 0;
  if (L == 0) { t.stage = 0; if (tr) *tr = t; return 0; }
  uint8_t *dsq = (uint8_t *)malloc((size_t)L + 2);
  dsq[0] = 255;
  memcpy(dsq + 1, seq, (size_t)L);
  dsq[L + 1] = 255;
  double F1 = o->do_max ? 1.0 : o->F1, F2 = o->do_max ? 1.0 : o->F2, F3 = o->do_max ? 1.0 : o->F3;
  int do_bias = o->do_max ? 0 : o->do_bias;
  float nullsc = orc_null1(L), usc, filtersc, vfsc = 0, fwdsc = 0, seq_score;
  double P;
  t.nullsc = nullsc;
  if (o->simd_msv == 0 || orc_msv_simd(p, dsq, L, &usc, &t.msv_xJ) < 0) orc_msv(p, dsq, L, &usc, &t.msv_xJ);
  t.usc = usc;
  seq_score = (float)((usc - nullsc) / LN2);
  P = orc_gumbel_surv(seq_score, ev[0], ev[1]);
  t.stage = 0;
  if (P > F1) goto done;
  if (do_bias) {
    filtersc = orc_bias_filtersc(p, dsq, L);
    seq_score = (float)((usc - filtersc) / LN2);
    P = orc_gumbel_surv(seq_score, ev[0], ev[1]);
    t.filtersc = filtersc;
    t.stage = 1;
    if (P > F1) goto done;
  } else
    filtersc = nullsc;
  t.filtersc = filtersc;
  t.stage = 2;
  if (P > F2) {
    if (o->simd_vit == 0 || orc_vit_simd(p, dsq, L, &vfsc, &t.vit_xC) < 0) orc_vit(p, dsq, L, &vfsc, &t.vit_xC);
    t.vfsc = vfsc;
    seq_score = (float)((vfsc - filtersc) / LN2);
    P = orc_gumbel_surv(seq_score, ev[2], ev[3]);
    if (P > F2) goto done;
  }
  t.stage = 3;
  fwdsc = orc_fwd_parser(p, dsq, L, L, 1, &t.fwd_exp, &t.fwd_mant);
  t.fwdsc = fwdsc;
  seq_score = (float)((fwdsc - filtersc) / LN2);
  P = orc_exp_surv(seq_score, ev[4], ev[5]);
  if (P > F3) goto done;
  t.stage = 4;
  {
    P7D_DOMAIN doms[MAXDOM_PER_SEQ];
    float *n2sc = (float *)calloc((size_t)L + 2, sizeof(float));
    int nregions, nmulti;
    int ndom = p7d_define_domains(p, dsq, L, o->do_null2, n2sc, doms, MAXDOM_PER_SEQ, model, target, poolcap,
                                  &nregions, &nmulti);
    if (ndom > 0) {
      t.stage = 5;
      float seqbias, pre_score, sum_score, pre2_score;
      const float omega = 1.0f / 256.0f;
      if (o->do_null2) {
        seqbias = 0.0f;
        for (int i = 0; i <= L; i++) seqbias += n2sc[i]; /* esl_vec_FSum over n2sc[0..L] */
        seqbias = orc_flogsum(0.0f, (float)(log((double)omega) + seqbias));
      } else
        seqbias = 0.0f;
      pre_score = (float)((fwdsc - nullsc) / LN2);
      seq_score = (float)((fwdsc - (nullsc + seqbias)) / LN2);
      float seqbias_used = seqbias;
      sum_score = 0.0f;
      float sbias = 0.0f;
      int Ld = 0;
      if (o->do_null2) {
        for (int d = 0; d < ndom; d++)
          if (doms[d].envsc - doms[d].domcorrection > 0.0f) {
            sum_score += doms[d].envsc;
            Ld += doms[d].jenv - doms[d].ienv + 1;
            sbias += doms[d].domcorrection;
          }
        sbias = orc_flogsum(0.0f, (float)(log((double)omega) + sbias));
      } else {
        for (int d = 0; d < ndom; d++)
          if (doms[d].envsc > 0.0f) {
            sum_score += doms[d].envsc;
            Ld += doms[d].jenv - doms[d].ienv + 1;
          }
        sbias = 0.0f;
      }
      sum_score += (float)((L - Ld) * log((double)((float)L / (float)(L + 3))));
      pre2_score = (float)((sum_score - nullsc) / LN2);
      sum_score = (float)((sum_score - (nullsc + sbias)) / LN2);
      if (Ld > 0 && sum_score > seq_score) {
        seq_score = sum_score;
        pre_score = pre2_score;
        seqbias_used = sbias;
      }
      double lnP = orc_exp_logsurv(seq_score, ev[4], ev[5]);
      for (int d = 0; d < ndom && nh < maxout; d++) {
        ORC_HIT *h = &out[nh++];
        memset(h, 0, sizeof(*h));
        int Ldd = doms[d].jenv - doms[d].ienv + 1;
        float bits = doms[d].envsc + (float)((L - Ldd) * log((double)((float)L / (float)(L + 3))));
        float dombias = o->do_null2 ? orc_flogsum(0.0f, (float)(log((double)omega) + doms[d].domcorrection)) : 0.0f;
        bits = (float)((bits - (nullsc + dombias)) / LN2);
        h->target = tidx;
        h->ndom_in_seq = ndom;
        h->dom_idx = d;
        h->ienv = doms[d].ienv; h->jenv = doms[d].jenv;
        h->iali = doms[d].iali; h->jali = doms[d].jali;
        h->hmmfrom = doms[d].hmmfrom; h->hmmto = doms[d].hmmto;
        h->envsc = doms[d].envsc; h->domcorrection = doms[d].domcorrection; h->oasc = doms[d].oasc;
        h->dombias = dombias; h->dom_bits = bits;
        h->dom_lnP = orc_exp_logsurv(bits, ev[4], ev[5]);
        h->seq_bits = seq_score; h->pre_bits = pre_score; h->seqbias = seqbias_used;
        h->fwdsc = fwdsc; h->nullsc = nullsc; h->seq_lnP = lnP;
        h->ali_off = doms[d].ali_off; h->ali_len = doms[d].ali_len;
        h->multidomain_region = doms[d].multidomain_region;
        *pool_used = doms[d].ali_off + 2 * (doms[d].ali_len + 1);
      }
    }
    free(n2sc);
  }
done:
  free(dsq);
  if (tr) *tr = t;
  return nh;
}

int orc_search(const ORC_PROFILE *p, const uint8_t *dsq, const int64_t *off, int n, const ORC_OPTS *o,
               ORC_HIT *hits, int maxhits, char *model_pool, char *target_pool, int poolcap, ORC_TRACE *trace,
               int nthreads) {
  int nhits = 0, used = 0, overflow = 0;
  orc_flogsum(0.0f, 0.0f); /* init table before threads */
  if (nthreads < 1) nthreads = 1;
  /* per-target hit counts first, then stable placement in target order */
  int *cnt = (int *)calloc((size_t)n + 1, sizeof(int));
  ORC_HIT **loc = (ORC_HIT **)calloc((size_t)n, sizeof(ORC_HIT *));
  char **lm = (char **)calloc((size_t)n, sizeof(char *)), **lt = (char **)calloc((size_t)n, sizeof(char *));
  int *lused = (int *)calloc((size_t)n, sizeof(int));
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
  for (int t = 0; t < n; t++) {
    int L = (int)(off[t + 1] - off[t]);
    int cap = 8 * L + 8 * p->M + 128;
    ORC_HIT tmp[MAXDOM_PER_SEQ];
    char *m = (char *)malloc((size_t)cap), *g = (char *)malloc((size_t)cap);
    int pu = 0;
    int nh = one_target(p, dsq + off[t], L, t, o, trace ? &trace[t] : NULL, tmp, MAXDOM_PER_SEQ, m, g, cap, &pu);
    if (nh > 0) {
      loc[t] = (ORC_HIT *)malloc(sizeof(ORC_HIT) * nh);
      memcpy(loc[t], tmp, sizeof(ORC_HIT) * nh);
      lm[t] = m;
      lt[t] = g;
      lused[t] = pu;
      cnt[t] = nh;
    } else {
      free(m);
      free(g);
    }
  }
  for (int t = 0; t < n; t++) {
    if (!cnt[t]) continue;
    if (nhits + cnt[t] > maxhits || used + lused[t] > poolcap) overflow = 1;
    if (!overflow) {
      memcpy(model_pool + used, lm[t], (size_t)lused[t]);
      memcpy(target_pool + used, lt[t], (size_t)lused[t]);
      for (int d = 0; d < cnt[t]; d++) {
        hits[nhits] = loc[t][d];
        hits[nhits].ali_off += used;
        nhits++;
      }
      used += lused[t];
    }
    free(loc[t]);
    free(lm[t]);
    free(lt[t]);
  }
  free(cnt);
  free(loc);
  free(lm);
  free(lt);
  free(lused);
  return overflow ? -1 : nhits;
}

long long orc_search_many(const ORC_PROFILE *const *profs, int nprof, const uint8_t *dsq, const int64_t *off, int n,
                          const ORC_OPTS *o, int nthreads, long long *stage_counts) {
  orc_flogsum(0.0f, 0.0f); /* init table before threads */
  if (nthreads < 1) nthreads = 1;
  const int BLK = 1024;
  const int nblk = (n + BLK - 1) / BLK;
  int maxL = 0, maxM = 0;
  for (int t = 0; t < n; t++)
    if ((int)(off[t + 1] - off[t]) > maxL) maxL = (int)(off[t + 1] - off[t]);
  for (int h = 0; h < nprof; h++)
    if (profs[h]->M > maxM) maxM = profs[h]->M;
  const int cap = 8 * maxL + 8 * maxM + 128;
  long long ndom = 0, st[6] = {0, 0, 0, 0, 0, 0};
  int failed = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : ndom, st[:6])
  {
    char *m = (char *)malloc((size_t)cap), *g = (char *)malloc((size_t)cap);
    ORC_HIT tmp[MAXDOM_PER_SEQ];
    /* profile-major: the threads walk the blocks of one profile together, so its tables stay in the shared cache */
#pragma omp for schedule(dynamic, 1)
    for (long long w = 0; w < (long long)nprof * nblk; w++) {
      const int h = (int)(w / nblk), b = (int)(w % nblk);
      const int t1 = (b + 1) * BLK < n ? (b + 1) * BLK : n;
      for (int t = b * BLK; t < t1; t++) {
        ORC_TRACE tr;
        int pu = 0;
        const int nh = one_target(profs[h], dsq + off[t], (int)(off[t + 1] - off[t]), t, o, &tr, tmp, MAXDOM_PER_SEQ, m, g, cap, &pu);
        if (nh < 0) failed = 1;
        else ndom += nh;
        st[tr.stage]++;
      }
    }
    free(m);
    free(g);
  }
  if (stage_counts)
    for (int i = 0; i < 6; i++) stage_counts[i] = st[i];
  return failed ? -1 : ndom;
}

long long orc_translate_reads(const char *nt, const int64_t *off, int nreads, int table, uint8_t *dsq_out, int64_t *off_out,
                              int nthreads) {
  if (nthreads < 1) nthreads = 1;
  int64_t pos = 0;
  for (int r = 0; r < nreads; r++) {
    const int L = (int)(off[r + 1] - off[r]);
    for (int f = 0; f < 6; f++) {
      const int j = f % 3;
      off_out[(size_t)6 * r + f] = pos;
      pos += (L - j) >= 3 ? (L - j) / 3 : 0;
    }
  }
  off_out[(size_t)6 * nreads] = pos;
  int bad = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int r = 0; r < nreads; r++) {
    const int L = (int)(off[r + 1] - off[r]);
    char *rc = (char *)malloc((size_t)L + 1), *aa = (char *)malloc((size_t)L / 3 + 2);
    orc_revcomp(nt + off[r], L, rc);
    for (int f = 0; f < 6; f++) {
      const int j = f % 3;
      const char *src = (f < 3 ? nt + off[r] : rc) + j;
      const int naa = (L - j) >= 3 ? orc_translate(src, L - j, table, aa) : 0;
      if (naa < 0) { bad = 1; continue; }
      orc_digitize(ORC_AMINO, aa, naa, dsq_out + off_out[(size_t)6 * r + f]);
    }
    free(rc);
    free(aa);
  }
  return bad ? -1 : (long long)6 * nreads;
}
