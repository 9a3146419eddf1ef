/*
 * p7_core.c -- ORACLE (test infrastructure): alphabets, HMMER3/f reader, profile
 * configuration and the integer/float "optimized profile" tables.
 * Restates SURVEY.md Appendix A.2-A.5 (HMMER3 is not vendored in the reference;
 * call site: /root/reference/lib/aln.py:664-667).
 */
#include "p7oracle.h"
#include "p7_internal.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#pragma STDC FP_CONTRACT OFF

static const char *AMINO_SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";
static const char *DNA_SYMS = "ACGT-RYMKSWHBVDN*~";

/* Swiss-Prot 50.8 background (A.3) */
static const float AMINO_BG[20] = {
    0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f,
    0.0590092f, 0.0594422f, 0.0963728f, 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f,
    0.0540978f, 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

static int abc_K(int abc) { return abc == ORC_AMINO ? 20 : 4; }
static int abc_Kp(int abc) { return abc == ORC_AMINO ? 29 : 18; }
static const char *abc_syms(int abc) { return abc == ORC_AMINO ? AMINO_SYMS : DNA_SYMS; }

/* degeneracy: fills mem[K] with 0/1 for code x (K < x <= Kp-3) */
static void degen_members(int abc, int x, int *mem) {
  int K = abc_K(abc);
  const char *syms = abc_syms(abc);
  const char *set = "";
  memset(mem, 0, sizeof(int) * K);
  if (abc == ORC_AMINO) {
    switch (syms[x]) {
      case 'B': set = "DN"; break;
      case 'J': set = "IL"; break;
      case 'Z': set = "EQ"; break;
      case 'O': set = "K"; break;
      case 'U': set = "C"; break;
      case 'X': set = "ACDEFGHIKLMNPQRSTVWY"; break;
    }
  } else {
    switch (syms[x]) {
      case 'R': set = "AG"; break;
      case 'Y': set = "CT"; break;
      case 'M': set = "AC"; break;
      case 'K': set = "GT"; break;
      case 'S': set = "CG"; break;
      case 'W': set = "AT"; break;
      case 'H': set = "ACT"; break;
      case 'B': set = "CGT"; break;
      case 'V': set = "ACG"; break;
      case 'D': set = "AGT"; break;
      case 'N': set = "ACGT"; break;
    }
  }
  for (const char *c = set; *c; c++) {
    const char *q = strchr(syms, *c);
    mem[q - syms] = 1;
  }
}

char orc_sym(int abc, int code) { return abc_syms(abc)[code]; }

int orc_digitize(int abc, const char *seq, int n, uint8_t *out) {
  const char *syms = abc_syms(abc);
  int bad = 0;
  for (int i = 0; i < n; i++) {
    char c = (char)toupper((unsigned char)seq[i]);
    if (abc == ORC_DNA && c == 'U') c = 'T';
    if (abc == ORC_DNA && c == 'X') c = 'N';
    const char *q = (c == 0) ? NULL : strchr(syms, c);
    if (!q) {
      bad++;
      out[i] = (uint8_t)(abc_Kp(abc) - 3); /* any */
    } else
      out[i] = (uint8_t)(q - syms);
  }
  return bad;
}

/* ---------------- HMM file ---------------- */

static float nlp2p(const char *tok) {
  if (tok[0] == '*') return 0.0f;
  return expf((float)(-1.0 * atof(tok)));
}

static ORC_HMM *hmm_alloc(int M, int abc) {
  ORC_HMM *h = (ORC_HMM *)calloc(1, sizeof(ORC_HMM));
  h->M = M;
  h->abc = abc;
  h->K = abc_K(abc);
  h->Kp = abc_Kp(abc);
  h->mat = (float *)calloc((size_t)(M + 1) * h->K, sizeof(float));
  h->ins = (float *)calloc((size_t)(M + 1) * h->K, sizeof(float));
  h->t = (float *)calloc((size_t)(M + 1) * 7, sizeof(float));
  h->cons = (char *)calloc((size_t)M + 2, 1);
  return h;
}

static void hmm_default_consensus(ORC_HMM *h) {
  /* p7_hmm_SetConsensus: argmax emission; upper-case if p >= 0.5 (amino) / 0.9 (DNA) */
  float thr = h->abc == ORC_AMINO ? 0.5f : 0.9f;
  for (int k = 1; k <= h->M; k++) {
    int best = 0;
    for (int x = 1; x < h->K; x++)
      if (h->mat[k * h->K + x] > h->mat[k * h->K + best]) best = x;
    char c = abc_syms(h->abc)[best];
    h->cons[k] = (h->mat[k * h->K + best] >= thr) ? c : (char)tolower((unsigned char)c);
  }
}

static void hmm_default_compo(ORC_HMM *h) {
  /* p7_hmm_SetComposition is what hmmbuild writes; when absent fall back to the
   * unweighted mean match emission (documented deviation, only used by the bias filter) */
  for (int x = 0; x < h->K; x++) {
    double s = 0;
    for (int k = 1; k <= h->M; k++) s += h->mat[k * h->K + x];
    h->compo[x] = (float)(s / h->M);
  }
  h->has_compo = 1;
}

void orc_hmm_free(ORC_HMM *h) {
  if (!h) return;
  free(h->mat);
  free(h->ins);
  free(h->t);
  free(h->cons);
  free(h);
}
int orc_hmm_M(const ORC_HMM *h) { return h->M; }
int orc_hmm_abc(const ORC_HMM *h) { return h->abc; }
const char *orc_hmm_name(const ORC_HMM *h) { return h->name; }
void orc_hmm_set_stats(ORC_HMM *h, const float *ev) {
  memcpy(h->ev, ev, sizeof(float) * 6);
  h->has_stats = 1;
}
void orc_hmm_get_stats(const ORC_HMM *h, float *ev) { memcpy(ev, h->ev, sizeof(float) * 6); }

#define FAIL(...)                         \
  do {                                    \
    if (err) snprintf(err, errlen, __VA_ARGS__); \
    if (fp) fclose(fp);                   \
    orc_hmm_free(h);                      \
    free(line);                           \
    return NULL;                          \
  } while (0)

static int next_tokens(char *line, char **tok, int maxtok) {
  int n = 0;
  char *save = NULL;
  for (char *t = strtok_r(line, " \t\r\n", &save); t && n < maxtok; t = strtok_r(NULL, " \t\r\n", &save))
    tok[n++] = t;
  return n;
}

ORC_HMM *orc_hmm_read(const char *path, char *err, int errlen) {
  FILE *fp = fopen(path, "r");
  ORC_HMM *h = NULL;
  size_t cap = 1 << 16;
  char *line = (char *)malloc(cap);
  char *tok[64];
  int M = -1, abc = -1, cons_flag = 0, map_flag = 0;
  char name[256] = "";
  float ev[6];
  int got[3] = {0, 0, 0};
  if (!fp) FAIL("cannot open %s", path);
  if (!fgets(line, (int)cap, fp) || strncmp(line, "HMMER3/", 7) != 0) FAIL("%s: not a HMMER3 ASCII file", path);
  /* header */
  for (;;) {
    if (!fgets(line, (int)cap, fp)) FAIL("%s: premature end of header", path);
    if (strncmp(line, "HMM ", 4) == 0 || strncmp(line, "HMM\t", 4) == 0) break;
    char copy[1024];
    strncpy(copy, line, sizeof(copy) - 1);
    copy[sizeof(copy) - 1] = 0;
    int nt = next_tokens(copy, tok, 16);
    if (nt < 2) continue;
    if (!strcmp(tok[0], "NAME")) strncpy(name, tok[1], 255);
    else if (!strcmp(tok[0], "LENG")) M = atoi(tok[1]);
    else if (!strcmp(tok[0], "ALPH")) {
      if (!strcasecmp(tok[1], "amino")) abc = ORC_AMINO;
      else if (!strcasecmp(tok[1], "DNA") || !strcasecmp(tok[1], "RNA")) abc = ORC_DNA;
      else FAIL("%s: unsupported alphabet %s", path, tok[1]);
    } else if (!strcmp(tok[0], "CONS")) cons_flag = !strcasecmp(tok[1], "yes");
    else if (!strcmp(tok[0], "MAP")) map_flag = !strcasecmp(tok[1], "yes");
    else if (!strcmp(tok[0], "STATS") && nt >= 5) {
      int w = -1;
      if (!strcmp(tok[2], "MSV")) w = 0;
      else if (!strcmp(tok[2], "VITERBI")) w = 1;
      else if (!strcmp(tok[2], "FORWARD")) w = 2;
      if (w >= 0) {
        ev[2 * w] = (float)atof(tok[3]);
        ev[2 * w + 1] = (float)atof(tok[4]);
        got[w] = 1;
      }
    }
  }
  (void)map_flag;
  if (M < 1 || abc < 0) FAIL("%s: missing LENG/ALPH", path);
  h = hmm_alloc(M, abc);
  snprintf(h->name, sizeof(h->name), "%s", name);
  if (got[0] && got[1] && got[2]) {
    memcpy(h->ev, ev, sizeof(ev));
    h->has_stats = 1;
  }
  int K = h->K;
  /* transition label line */
  if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated", path);
  /* optional COMPO, then node-0 insert line */
  if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated", path);
  int nt = next_tokens(line, tok, 64);
  if (nt > 0 && !strcmp(tok[0], "COMPO")) {
    if (nt < K + 1) FAIL("%s: short COMPO", path);
    for (int x = 0; x < K; x++) h->compo[x] = nlp2p(tok[x + 1]);
    h->has_compo = 1;
    if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated", path);
    nt = next_tokens(line, tok, 64);
  }
  if (nt < K) FAIL("%s: short node-0 insert line", path);
  for (int x = 0; x < K; x++) h->ins[x] = nlp2p(tok[x]);
  if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated", path);
  nt = next_tokens(line, tok, 64);
  if (nt < 7) FAIL("%s: short node-0 transition line", path);
  for (int x = 0; x < 7; x++) h->t[x] = nlp2p(tok[x]);
  h->mat[0] = 1.0f;
  for (int k = 1; k <= M; k++) {
    if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated at node %d", path, k);
    nt = next_tokens(line, tok, 64);
    if (nt < K + 1 || atoi(tok[0]) != k) FAIL("%s: bad match line at node %d", path, k);
    for (int x = 0; x < K; x++) h->mat[k * K + x] = nlp2p(tok[x + 1]);
    if (cons_flag && nt >= K + 3) h->cons[k] = tok[K + 2][0];
    if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated at node %d", path, k);
    nt = next_tokens(line, tok, 64);
    if (nt < K) FAIL("%s: bad insert line at node %d", path, k);
    for (int x = 0; x < K; x++) h->ins[k * K + x] = nlp2p(tok[x]);
    if (!fgets(line, (int)cap, fp)) FAIL("%s: truncated at node %d", path, k);
    nt = next_tokens(line, tok, 64);
    if (nt < 7) FAIL("%s: bad transition line at node %d", path, k);
    for (int x = 0; x < 7; x++) h->t[k * 7 + x] = nlp2p(tok[x]);
  }
  if (!cons_flag) hmm_default_consensus(h);
  if (!h->has_compo) hmm_default_compo(h);
  fclose(fp);
  free(line);
  return h;
}

ORC_HMM *orc_hmm_from_arrays(int M, int abc, const double *mat_nlp, const double *ins_nlp,
                             const double *t_nlp, const double *compo_nlp, const char *cons,
                             const float *ev, const char *name) {
  ORC_HMM *h = hmm_alloc(M, abc);
  int K = h->K;
  /* same conversion as the text reader: p = expf((float)(-v)); inf -> 0 */
  for (int i = 0; i < (M + 1) * K; i++) {
    h->mat[i] = isinf(mat_nlp[i]) ? 0.0f : expf((float)(-1.0 * mat_nlp[i]));
    h->ins[i] = isinf(ins_nlp[i]) ? 0.0f : expf((float)(-1.0 * ins_nlp[i]));
  }
  for (int i = 0; i < (M + 1) * 7; i++) h->t[i] = isinf(t_nlp[i]) ? 0.0f : expf((float)(-1.0 * t_nlp[i]));
  if (compo_nlp) {
    for (int x = 0; x < K; x++) h->compo[x] = isinf(compo_nlp[x]) ? 0.0f : expf((float)(-1.0 * compo_nlp[x]));
    h->has_compo = 1;
  } else
    hmm_default_compo(h);
  if (cons) {
    for (int k = 1; k <= M; k++) h->cons[k] = cons[k - 1];
  } else
    hmm_default_consensus(h);
  if (ev) orc_hmm_set_stats(h, ev);
  if (name) strncpy(h->name, name, 255);
  return h;
}

/* ---------------- profile ---------------- */

static uint8_t unbiased_byteify(float scale_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.f) ? 255 : (uint8_t)sc;
}
static uint8_t biased_byteify(float scale_b, uint8_t bias_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > (float)(255 - bias_b)) ? 255 : (uint8_t)((uint8_t)sc + bias_b);
}
static int16_t wordify(float scale_w, float sc) {
  sc = roundf(scale_w * sc);
  if (sc >= 32767.0f) return 32767;
  if (sc <= -32768.0f) return -32768;
  return (int16_t)sc;
}

ORC_PROFILE *orc_profile_create(const ORC_HMM *h) {
  ORC_PROFILE *p = (ORC_PROFILE *)calloc(1, sizeof(ORC_PROFILE));
  int M = h->M, K = h->K, Kp = h->Kp;
  p->hmm = h;
  p->M = M;
  p->K = K;
  p->Kp = Kp;
  p->abc = h->abc;
  for (int x = 0; x < K; x++) p->bgf[x] = (h->abc == ORC_AMINO) ? AMINO_BG[x] : 0.25f;
  p->tsc = (float *)malloc(sizeof(float) * (M + 1) * 8);
  p->msc = (float *)malloc(sizeof(float) * Kp * (M + 1));
  p->rbv = (uint8_t *)malloc((size_t)Kp * (M + 1));
  p->Qb = (M + 31) / 32;
  p->rbv_s = NULL;
  p->rwv = (int16_t *)malloc(sizeof(int16_t) * Kp * (M + 1));
  p->twv = (int16_t *)malloc(sizeof(int16_t) * (M + 1) * 8);
  p->rfv = (float *)malloc(sizeof(float) * Kp * (M + 1));
  p->tfv = (float *)malloc(sizeof(float) * (M + 1) * 8);
  for (int i = 0; i < (M + 1) * 8; i++) p->tsc[i] = -INFINITY;

  /* occupancy -> local entry (A.3) */
  float *occ = (float *)calloc((size_t)M + 1, sizeof(float));
  occ[1] = h->t[ORC_MI] + h->t[ORC_MM];
  for (int k = 2; k <= M; k++)
    occ[k] = (float)(occ[k - 1] * (h->t[(k - 1) * 7 + ORC_MM] + h->t[(k - 1) * 7 + ORC_MI]) +
                     (1.0 - occ[k - 1]) * h->t[(k - 1) * 7 + ORC_DM]);
  float Z = 0.f;
  for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
  for (int k = 1; k <= M; k++) p->tsc[(k - 1) * 8 + ORC_BM] = (float)log((double)(occ[k] / Z));
  free(occ);
  for (int k = 1; k < M; k++)
    for (int t = 0; t < 7; t++) p->tsc[k * 8 + t] = (float)log((double)h->t[k * 7 + t]);
  /* D_1 is wing-retracted: transitions out of D_1 impossible */
  if (M > 1) {
    p->tsc[1 * 8 + ORC_DM] = -INFINITY;
    p->tsc[1 * 8 + ORC_DD] = -INFINITY;
  }

  /* match scores */
  float sc[32];
  int mem[20];
  for (int x = 0; x < Kp; x++) p->msc[x * (M + 1)] = -INFINITY;
  for (int k = 1; k <= M; k++) {
    for (int x = 0; x < K; x++) sc[x] = (float)log((double)(h->mat[k * K + x] / p->bgf[x]));
    sc[K] = -INFINITY;
    sc[Kp - 2] = -INFINITY;
    sc[Kp - 1] = -INFINITY;
    for (int x = K + 1; x <= Kp - 3; x++) {
      degen_members(h->abc, x, mem);
      float result = 0.f, denom = 0.f;
      for (int y = 0; y < K; y++)
        if (mem[y]) {
          result += sc[y] * p->bgf[y];
          denom += p->bgf[y];
        }
      sc[x] = result / denom;
    }
    for (int x = 0; x < Kp; x++) p->msc[x * (M + 1) + k] = sc[x];
  }

  /* MSV bytes (A.4) */
  p->scale_b = (float)(3.0 / 0.69314718055994529);
  p->base_b = 190;
  float mx = 0.0f; /* HMMER's max runs over match AND insert scores; inserts are 0 */
  for (int x = 0; x < K; x++)
    for (int k = 1; k <= M; k++)
      if (p->msc[x * (M + 1) + k] > mx) mx = p->msc[x * (M + 1) + k];
  p->bias_b = unbiased_byteify(p->scale_b, -1.0f * mx);
  for (int x = 0; x < Kp; x++) {
    p->rbv[x * (M + 1)] = 255;
    for (int k = 1; k <= M; k++) p->rbv[x * (M + 1) + k] = biased_byteify(p->scale_b, p->bias_b, p->msc[x * (M + 1) + k]);
  }
  p->tbm_b = unbiased_byteify(p->scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
  p->tec_b = unbiased_byteify(p->scale_b, logf(0.5f));
  /* striped copy for orc_msv_simd: cell k -> vector (k-1) % Qb, byte lane (k-1) / Qb; padding costs 255 */
  if (posix_memalign((void **)&p->rbv_s, 32, (size_t)Kp * p->Qb * 32) != 0) p->rbv_s = NULL;
  if (p->rbv_s) {
    memset(p->rbv_s, 255, (size_t)Kp * p->Qb * 32);
    for (int x = 0; x < Kp; x++)
      for (int k = 1; k <= M; k++) p->rbv_s[((size_t)x * p->Qb + (k - 1) % p->Qb) * 32 + (k - 1) / p->Qb] = p->rbv[x * (M + 1) + k];
  }

  /* Viterbi words (A.5) */
  p->scale_w = (float)(500.0 / 0.69314718055994529);
  p->base_w = 12000;
  for (int x = 0; x < Kp; x++) {
    p->rwv[x * (M + 1)] = -32768;
    for (int k = 1; k <= M; k++) p->rwv[x * (M + 1) + k] = wordify(p->scale_w, p->msc[x * (M + 1) + k]);
  }
  for (int k = 0; k <= M; k++)
    for (int t = 0; t < 8; t++) {
      int16_t v = wordify(p->scale_w, p->tsc[k * 8 + t]);
      int16_t cap = (t == ORC_II) ? -1 : 0;
      p->twv[k * 8 + t] = v < cap ? v : cap;
    }

  /* striped copies for orc_vit_simd */
  p->Qw = (M + 15) / 16;
  p->rwv_s = p->twv_s = NULL;
  if (posix_memalign((void **)&p->rwv_s, 32, (size_t)Kp * p->Qw * 32) != 0) p->rwv_s = NULL;
  if (posix_memalign((void **)&p->twv_s, 32, (size_t)8 * p->Qw * 32) != 0) p->twv_s = NULL;
  if (p->rwv_s && p->twv_s) {
    const int Qw = p->Qw;
    for (size_t i = 0; i < (size_t)Kp * Qw * 16; i++) p->rwv_s[i] = -32768;
    for (size_t i = 0; i < (size_t)8 * Qw * 16; i++) p->twv_s[i] = -32768;
    static const int kind[8] = {ORC_BM, ORC_MM, ORC_IM, ORC_DM, ORC_MI, ORC_II, ORC_MD, ORC_DD};
    for (int k = 1; k <= M; k++) {
      const size_t pos = (size_t)((k - 1) % Qw) * 16 + (size_t)((k - 1) / Qw);
      for (int x = 0; x < Kp; x++) p->rwv_s[(size_t)x * Qw * 16 + pos] = p->rwv[x * (M + 1) + k];
      for (int z = 0; z < 8; z++) {
        const int node = (z == 4 || z == 5) ? k : k - 1;   /* MI, II belong to node k; the others lead from k-1 to k */
        p->twv_s[(size_t)z * Qw * 16 + pos] = p->twv[node * 8 + kind[z]];
      }
    }
  }

  /* float odds */
  for (int i = 0; i < Kp * (M + 1); i++) p->rfv[i] = expf(p->msc[i]);
  for (int i = 0; i < (M + 1) * 8; i++) p->tfv[i] = expf(p->tsc[i]);
  return p;
}

void orc_profile_free(ORC_PROFILE *p) {
  if (!p) return;
  free(p->tsc);
  free(p->msc);
  free(p->rbv);
  free(p->rbv_s);
  free(p->rwv_s);
  free(p->twv_s);
  free(p->rwv);
  free(p->twv);
  free(p->rfv);
  free(p->tfv);
  free(p);
}

uint8_t p7i_unbiased_byteify(float scale_b, float sc) { return unbiased_byteify(scale_b, sc); }
int16_t p7i_wordify(float scale_w, float sc) { return wordify(scale_w, sc); }
void p7i_degen_members(int abc, int x, int *mem) { degen_members(abc, x, mem); }

/* bias-filter state-1 emission odds: compo/bg, degenerate = sum e / sum f over members,
 * gap, '*', '~' = 1 (esl_hmm_Configure) */
void p7i_bias_eo1(const ORC_PROFILE *p, float *eo1) {
  int K = p->K, Kp = p->Kp, mem[20];
  for (int x = 0; x < K; x++) eo1[x] = p->hmm->compo[x] / p->bgf[x];
  for (int x = K + 1; x <= Kp - 3; x++) {
    degen_members(p->abc, x, mem);
    float num = 0.0f, den = 0.0f;
    for (int y = 0; y < K; y++)
      if (mem[y]) {
        num += p->hmm->compo[y];
        den += p->bgf[y];
      }
    eo1[x] = num / den;
  }
  eo1[K] = eo1[Kp - 2] = eo1[Kp - 1] = 1.0f;
}

/* ---------------- statistics ---------------- */
double orc_gumbel_surv(double x, double mu, double lambda) {
  double y = lambda * (x - mu);
  double ey = -exp(-y);
  if (fabs(ey) < 5e-9) return -ey;
  return 1.0 - exp(ey);
}
double orc_exp_surv(double x, double mu, double lambda) {
  if (x < mu) return 1.0;
  return exp(-lambda * (x - mu));
}
double orc_exp_logsurv(double x, double mu, double lambda) {
  if (x < mu) return 0.0;
  return -lambda * (x - mu);
}

/* p7_FLogsum: table-driven log(e^a + e^b), 16000 entries at 1/1000 nat */
static float flogsum_tbl[16000];
static int flogsum_init = 0;
float orc_flogsum(float a, float b) {
  if (!flogsum_init) {
    for (int i = 0; i < 16000; i++) flogsum_tbl[i] = (float)log(1.0 + exp((double)-i / 1000.0));
    flogsum_init = 1;
  }
  float max = a > b ? a : b;
  float min = a > b ? b : a;
  return (min == -INFINITY || (max - min) >= 15.7f) ? max : max + flogsum_tbl[(int)((max - min) * 1000.0f)];
}

float orc_null1(int L) {
  float p1 = (float)L / (float)(L + 1);
  return (float)((double)L * log((double)p1) + log(1.0 - (double)p1));
}

/* accessors for tests (brute-force path enumeration against the configured profile) */
float orc_profile_tsc(const ORC_PROFILE *p, int k, int t) { return p->tsc[k * 8 + t]; }
float orc_profile_msc(const ORC_PROFILE *p, int x, int k) { return p->msc[x * (p->M + 1) + k]; }
int orc_profile_rbv(const ORC_PROFILE *p, int x, int k) { return p->rbv[x * (p->M + 1) + k]; }
int orc_profile_rwv(const ORC_PROFILE *p, int x, int k) { return p->rwv[x * (p->M + 1) + k]; }
int orc_profile_bytes(const ORC_PROFILE *p, int which) {
  switch (which) { case 0: return p->tbm_b; case 1: return p->tec_b; case 2: return p->bias_b; default: return p->base_b; }
}
