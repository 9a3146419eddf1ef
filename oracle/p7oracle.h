/*
 * p7oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar, obviously-correct restatement of the search hot path of PhyloAln
 * (reference: /root/reference/lib/aln.py:233-259 translation, :650-674 the
 * `hmmsearch` call, :482-570 hit extraction) together with the HMMER3
 * `hmmsearch` acceleration pipeline that the reference shells out to.
 *
 * HMMER itself (requirement "hmmer >=3.1", reference README.md:50) and
 * Biopython (>=1.77, README.md:49) are third-party dependencies that are NOT
 * vendored under /root/reference and are not installed in this image, so this
 * file restates their published algorithms (Eddy 2011, PLoS Comput Biol
 * 7:e1002195; HMMER 3.1-3.4 user guide).
 *
 *                  ***  PARITY UNPINNED at the HMMER boundary  ***
 * No golden hmmsearch outputs exist in the reference (tests/run_test.sh only
 * checks the exit status), and no hmmsearch binary is available to generate
 * them.  The PhyloAln-glue parts (translation frames, hit_info rows, column
 * mapping) ARE pinned against the reference's own Python, see tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (phyloaln_b200/)
 * never does.
 *
 * Floating-point stages use ONE canonical evaluation order (32-lane blocked
 * scans / butterfly sums, no FMA contraction, power-of-two rescaling) that
 * the CUDA kernels reproduce exactly; it is documented in DESIGN.md section
 * "Canonical float order".
 */
#ifndef P7ORACLE_H
#define P7ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NL 32              /* canonical lane count for blocked scans */
#define ORC_AMINO 0
#define ORC_DNA 1

enum { ORC_MM = 0, ORC_MI, ORC_MD, ORC_IM, ORC_II, ORC_DM, ORC_DD, ORC_BM, ORC_NT };

typedef struct {
  int M, abc, K, Kp;
  char name[256];
  float *mat;          /* (M+1)*K probabilities, node 0 unused              */
  float *ins;          /* (M+1)*K                                            */
  float *t;            /* (M+1)*7 probabilities MM MI MD IM II DM DD         */
  float compo[20];
  int has_compo;
  char *cons;          /* consensus letters, index 1..M                      */
  float ev[6];         /* MSV mu,lambda; VIT mu,lambda; FWD tau,lambda       */
  int has_stats;
} ORC_HMM;

typedef struct {
  const ORC_HMM *hmm;
  int M, K, Kp, abc;
  float bgf[20];
  float *tsc;          /* (M+1)*8 log transition scores, [k*8+t]; BM at k = B->M_{k+1} */
  float *msc;          /* Kp*(M+1) log-odds match scores [x*(M+1)+k]          */
  /* MSV 8-bit */
  uint8_t *rbv;        /* Kp*(M+1) biased costs                               */
  uint8_t *rbv_s;      /* striped copy for the AVX2 filter: Kp * Qb * 32 bytes (NULL until built) */
  int Qb;              /* vectors per row = ceil(M/32)                        */
  uint8_t tbm_b, tec_b, base_b, bias_b;
  float scale_b;
  /* Viterbi 16-bit */
  int16_t *rwv;        /* Kp*(M+1)                                            */
  int16_t *twv;        /* (M+1)*8                                             */
  float scale_w;
  int16_t base_w;
  /* float odds */
  float *rfv;          /* Kp*(M+1) = expf(msc)                                */
  float *tfv;          /* (M+1)*8  = expf(tsc)                                */
  /* striped copies for the AVX2 Viterbi filter (orc_vit_simd): cell k -> vector (k-1) % Qw, word lane (k-1) / Qw   */
  int16_t *rwv_s;      /* Kp * Qw * 16, padding -32768                        */
  int16_t *twv_s;      /* 8 * Qw * 16: BM MM IM DM (k-1 -> k) | MI II (k) | MD DD (k-1 -> k), padding -32768 */
  int Qw;              /* vectors per row = ceil(M/16)                        */
} ORC_PROFILE;

typedef struct {
  double F1, F2, F3;       /* 0.02 1e-3 1e-5 */
  double E, domE;          /* 10 10 */
  int do_bias, do_null2, do_max;
  double Z, domZ;          /* <=0: set by number of targets / reported */
  int simd_msv;            /* 1: MSV through the striped AVX2 filter (same numbers; bench.py's CPU baseline) */
  int simd_vit;            /* 1: Viterbi filter through the striped AVX2 filter (same numbers)                */
} ORC_OPTS;

typedef struct {
  int target;              /* index of the target sequence                      */
  int ndom_in_seq, dom_idx;
  int ienv, jenv, iali, jali, hmmfrom, hmmto;
  float envsc, domcorrection, oasc, dombias, dom_bits;
  double dom_lnP;
  float seq_bits, pre_bits, seqbias, fwdsc, nullsc;
  double seq_lnP;
  int ali_off, ali_len;    /* offset/len into the string pools                  */
  int multidomain_region;  /* 1 if the region tripped the rt3 test              */
} ORC_HIT;

/* per-target filter trace (debug / parity of raw scores) */
typedef struct {
  float nullsc, usc, filtersc, vfsc, fwdsc;
  int msv_xJ;              /* raw final xJ byte, -1 on overflow                 */
  int vit_xC;              /* raw final xC word, 32767 on overflow              */
  int fwd_exp;             /* power-of-two scale exponent of forward            */
  float fwd_mant;          /* xC*tCT mantissa                                   */
  int stage;               /* 0 fail msv,1 fail bias,2 fail vit,3 fail fwd,4 passed fwd, 5 produced domains */
} ORC_TRACE;

ORC_HMM *orc_hmm_read(const char *path, char *err, int errlen);
ORC_HMM *orc_hmm_from_arrays(int M, int abc, const double *mat_nlp, const double *ins_nlp,
                             const double *t_nlp, const double *compo_nlp, const char *cons,
                             const float *ev, const char *name);
void orc_hmm_free(ORC_HMM *h);
int orc_hmm_M(const ORC_HMM *h);
int orc_hmm_abc(const ORC_HMM *h);
void orc_hmm_set_stats(ORC_HMM *h, const float *ev);
void orc_hmm_get_stats(const ORC_HMM *h, float *ev);
const char *orc_hmm_name(const ORC_HMM *h);

ORC_PROFILE *orc_profile_create(const ORC_HMM *h);
void orc_profile_free(ORC_PROFILE *p);

float orc_profile_tsc(const ORC_PROFILE *p, int k, int t);
float orc_profile_msc(const ORC_PROFILE *p, int x, int k);
int orc_profile_rbv(const ORC_PROFILE *p, int x, int k);
/* Striped AVX2 MSV filter (Farrar layout, as HMMER's own SIMD filter): same result as orc_msv, bit for bit.
 * Returns -1 if the CPU has no AVX2 (callers then use orc_msv). */
int orc_msv_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xJ_raw);
int orc_have_avx2(void);
/* Striped AVX2 Viterbi filter (16 signed words per vector, Farrar layout with lazy D->D propagation, the way HMMER's
 * p7_ViterbiFilter is vectorised): same xC word and score as orc_vit, bit for bit. -1 without AVX2. */
int orc_vit_simd(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *ret_sc, int *xC_raw);
int orc_profile_rwv(const ORC_PROFILE *p, int x, int k);
int orc_profile_bytes(const ORC_PROFILE *p, int which);

int orc_digitize(int abc, const char *seq, int n, uint8_t *out);  /* returns #invalid */
char orc_sym(int abc, int code);

/* raw filters; dsq is 1-based: dsq[1..L] (dsq[0] ignored) */
float orc_null1(int L);
int orc_msv(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *sc, int *xJ_raw);
int orc_vit(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *sc, int *xC_raw);
float orc_bias_filtersc(const ORC_PROFILE *p, const uint8_t *dsq, int L);
float orc_fwd_parser(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit,
                     int *exp_out, float *mant_out);
float orc_bwd_parser_total(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit);
/* generic log-space Forward / Viterbi (textbook, double) for cross-checks */
double orc_generic_forward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit);
double orc_generic_viterbi(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit);

/* Full pipeline over a batch of digital targets.
 * dsq: concatenated residues (0-based, no sentinels); off[n+1] offsets.
 * hits: caller array of capacity maxhits; model/target pools of poolcap bytes.
 * returns number of hits (>=0) or -1 if capacity exceeded. Thresholds with Z/domZ
 * are NOT applied here (see orc_apply_thresholds).  trace may be NULL. */
int orc_search(const ORC_PROFILE *p, const uint8_t *dsq, const int64_t *off, int n,
               const ORC_OPTS *o, ORC_HIT *hits, int maxhits, char *model_pool, char *target_pool,
               int poolcap, ORC_TRACE *trace, int nthreads);

/* CPU baseline driver (bench.py): every profile against every target in ONE OpenMP region (profile-major blocks of
 * targets, per-thread scratch), hits counted but not returned.  stage_counts[6]: targets that ended in each ORC_TRACE.stage
 * summed over profiles.  Returns the number of domains, or -1. */
long long orc_search_many(const ORC_PROFILE *const *profs, int nprof, const uint8_t *dsq, const int64_t *off, int n,
                          const ORC_OPTS *o, int nthreads, long long *stage_counts);
/* Six-frame translation + digitisation of a chunk of reads (reference lib/aln.py:249-258 frame order: pos1..3, rev_pos1..3)
 * into the layout orc_search* takes: dsq_out (capacity 2*total nt + 8), off_out[6*nreads + 1].  Returns 6*nreads, -1 on an
 * invalid nucleotide. */
long long orc_translate_reads(const char *nt, const int64_t *off, int nreads, int table, uint8_t *dsq_out, int64_t *off_out,
                              int nthreads);

/* domain-definition internals exposed for tests */
int orc_domain_decoding(const ORC_PROFILE *p, const uint8_t *dsq, int L, float *btot, float *etot,
                        float *mocc);

/* translation (Biopython semantics, reference lib/aln.py:233-259) */
int orc_translate(const char *nt, int n, int table, char *out);       /* returns #aa or -1 invalid */
void orc_revcomp(const char *nt, int n, char *out);
int orc_codon_table(int table, char *out64);

double orc_gumbel_surv(double x, double mu, double lambda);
double orc_exp_surv(double x, double mu, double lambda);
double orc_exp_logsurv(double x, double mu, double lambda);
float orc_flogsum(float a, float b);

#ifdef __cplusplus
}
#endif
#endif
