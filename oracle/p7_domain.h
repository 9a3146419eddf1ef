/* p7_domain.h -- ORACLE internals (test infrastructure, not product code). */
#ifndef P7_DOMAIN_H
#define P7_DOMAIN_H
#include "p7oracle.h"
typedef struct {
  int ienv, jenv, iali, jali, hmmfrom, hmmto;
  float envsc, oasc, domcorrection;
  int ali_off, ali_len, multidomain_region;
} P7D_DOMAIN;
int p7d_rescore_envelope(const ORC_PROFILE *p, const uint8_t *dsq, int L, int ienv, int jenv, int do_null2,
                         float *n2sc, P7D_DOMAIN *dom, char *model, char *target, int alicap);
#define P7D_MAXENV 16 /* envelopes kept per multi-domain region */
int p7d_resolve_multidomain(const ORC_PROFILE *p, const uint8_t *dsq, int L, int ireg, int jreg, int *env_i, int *env_j,
                            int maxenv);
int p7d_define_domains(const ORC_PROFILE *p, const uint8_t *dsq, int L, int do_null2, float *n2sc,
                       P7D_DOMAIN *doms, int maxdom, char *model, char *target, int poolcap, int *nregions,
                       int *nmulti);
#endif
