/* p7_internal.h -- ORACLE internals (test infrastructure, not product code). */
#ifndef P7_INTERNAL_H
#define P7_INTERNAL_H
#include "p7oracle.h"
#include <stddef.h>

typedef struct { float tNN, tNB, tJJ, tJB, tCC, tCT, tEJ, tEC; } P7I_XF;
typedef struct { float xE, xN, xJ, xB, xC; } P7I_SPECIALS;
typedef struct { float *M, *I, *D; } P7I_MX; /* (L+1) rows of width M+2 */

uint8_t p7i_unbiased_byteify(float scale_b, float sc);
int16_t p7i_wordify(float scale_w, float sc);
void p7i_bias_eo1(const ORC_PROFILE *p, float *eo1);
void p7i_degen_members(int abc, int x, int *mem);

P7I_XF p7i_xf_config(int Lcfg, int multihit);
float p7i_lane_sum(const float *lane);
float p7i_fwd_row(const ORC_PROFILE *p, int x, const float *Mp, const float *Ip, const float *Dp, float xBp,
                  float *Mc, float *Ic, float *Dc, float *wA, float *wP);
void p7i_bwd_cells(const ORC_PROFILE *p, const float *we, const float *In, float xE, float *Mc, float *Ic,
                   float *Dc, float *wA, float *wP);
void p7i_bwd_row(const ORC_PROFILE *p, int x, const float *Mn, const float *In, const P7I_XF *xf,
                 const P7I_SPECIALS *sn, P7I_SPECIALS *sp, float *Mc, float *Ic, float *Dc, float *we,
                 float *wA, float *wP);
float p7i_forward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, P7I_MX *mx,
                  P7I_SPECIALS *sp, int *sexp, int *exp_out, float *mant_out);
float p7i_backward(const ORC_PROFILE *p, const uint8_t *dsq, int L, int Lcfg, int multihit, P7I_MX *mx,
                   P7I_SPECIALS *sp, int *sexp);
#endif
