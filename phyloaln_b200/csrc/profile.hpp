// profile.hpp -- host-side profile configuration for the B200 search path.
//
// Turns the numbers of one HMMER3/f profile (as written by hmmbuild, reference
// /root/reference/lib/aln.py:66-76) into the local multihit search profile and its three
// score systems: unsigned-byte MSV costs, signed-word Viterbi
// scores, float odds ratios for Forward/Backward.  Pure host C++ (no CUDA); the tables are
// re-laid-out for the kernels in api.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pa {

enum { T_MM = 0, T_MI, T_MD, T_IM, T_II, T_DM, T_DD, T_BM, T_N };

struct Alphabet {
  int type;  // 0 amino, 1 dna
  int K, Kp;
  const char *syms;
  static const Alphabet &get(int type);
  // members[x] = bitmask over K canonical residues for code x
  uint32_t members(int x) const;
};

struct HostProfile {
  std::string name;
  int abc = 0, M = 0, K = 0, Kp = 0;
  std::vector<float> mat, ins, t;  // probabilities, (M+1)*K / (M+1)*7
  std::vector<float> compo;        // K
  std::string cons;                // index 1..M (cons[0] unused)
  float ev[6] = {0, 0, 0, 0, 0, 0};
  bool has_stats = false;
  bool configured = false;   // configure() has run (done lazily, in parallel, by the commit)

  // configured profile (natural layouts, index [x*(M+1)+k] and [k*8+t])
  std::vector<float> tsc, msc, rfv, tfv, bgf;
  std::vector<uint8_t> rbv;
  std::vector<int16_t> rwv, twv;
  uint8_t tbm_b = 0, tec_b = 0, base_b = 190, bias_b = 0;
  float scale_b = 0, scale_w = 0;
  int16_t base_w = 12000;
  int16_t xw_E = 0;        // E->C / E->J word (multihit)
  float eo1[32];           // bias-filter state-1 emission odds per code

  void from_file_numbers(const char *name, int abc, int M, const double *mat_nlp, const double *ins_nlp,
                         const double *t_nlp, const double *compo_nlp, const char *cons, const float *ev);
  void configure();
};

uint8_t unbiased_byteify(float scale_b, float sc);
int16_t wordify(float scale_w, float sc);
float null1_score(int L);

// smallest float score s such that surv(s) <= F, found on the float grid by bisection
float gumbel_threshold(double F, double mu, double lambda);
float exp_threshold(double F, double mu, double lambda);
double gumbel_surv(double x, double mu, double lambda);
double exp_surv(double x, double mu, double lambda);
double exp_logsurv(double x, double mu, double lambda);
float flogsum(float a, float b);

}  // namespace pa
