// profile.cpp -- see profile.hpp. Host-only; compiled without FMA contraction so the
// quantised tables are reproducible (SURVEY.md section 7 "hard parts").
#include "profile.hpp"

#include <cmath>
#include <cstring>
#include <limits>

namespace pa {

static const float kAminoBg[20] = {0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f,
                                   0.0590092f, 0.0594422f, 0.0963728f, 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f,
                                   0.0540978f, 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

const Alphabet &Alphabet::get(int type) {
  static const Alphabet amino{0, 20, 29, "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"};
  static const Alphabet dna{1, 4, 18, "ACGT-RYMKSWHBVDN*~"};
  return type == 0 ? amino : dna;
}

uint32_t Alphabet::members(int x) const {
  if (x < K) return 1u << x;
  auto bit = [&](const char *set) {
    uint32_t m = 0;
    for (const char *c = set; *c; ++c) m |= 1u << (int)(std::strchr(syms, *c) - syms);
    return m;
  };
  char s = syms[x];
  if (type == 0) {
    switch (s) {
      case 'B': return bit("DN");
      case 'J': return bit("IL");
      case 'Z': return bit("EQ");
      case 'O': return bit("K");
      case 'U': return bit("C");
      case 'X': return (1u << 20) - 1;
    }
  } else {
    switch (s) {
      case 'R': return bit("AG");
      case 'Y': return bit("CT");
      case 'M': return bit("AC");
      case 'K': return bit("GT");
      case 'S': return bit("CG");
      case 'W': return bit("AT");
      case 'H': return bit("ACT");
      case 'B': return bit("CGT");
      case 'V': return bit("ACG");
      case 'D': return bit("AGT");
      case 'N': return 15u;
    }
  }
  return 0;  // gap, '*', '~'
}

uint8_t unbiased_byteify(float scale_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.f) ? 255 : (uint8_t)sc;
}
static uint8_t biased_byteify(float scale_b, uint8_t bias_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > (float)(255 - bias_b)) ? 255 : (uint8_t)((uint8_t)sc + bias_b);
}
int16_t wordify(float scale_w, float sc) {
  sc = roundf(scale_w * sc);
  if (sc >= 32767.0f) return 32767;
  if (sc <= -32768.0f) return -32768;
  return (int16_t)sc;
}

float null1_score(int L) {
  float p1 = (float)L / (float)(L + 1);
  return (float)((double)L * std::log((double)p1) + std::log(1.0 - (double)p1));
}

static float nlp2p(double v) { return std::isinf(v) ? 0.0f : expf((float)(-1.0 * v)); }

void HostProfile::from_file_numbers(const char *nm, int abc_, int M_, const double *mat_nlp, const double *ins_nlp,
                                    const double *t_nlp, const double *compo_nlp, const char *cons_, const float *ev_) {
  const Alphabet &a = Alphabet::get(abc_);
  name = nm ? nm : "";
  abc = abc_;
  M = M_;
  K = a.K;
  Kp = a.Kp;
  mat.resize((size_t)(M + 1) * K);
  ins.resize((size_t)(M + 1) * K);
  t.resize((size_t)(M + 1) * 7);
  for (size_t i = 0; i < mat.size(); i++) {
    mat[i] = nlp2p(mat_nlp[i]);
    ins[i] = nlp2p(ins_nlp[i]);
  }
  for (size_t i = 0; i < t.size(); i++) t[i] = nlp2p(t_nlp[i]);
  compo.assign(K, 0.0f);
  if (compo_nlp) {
    for (int x = 0; x < K; x++) compo[x] = nlp2p(compo_nlp[x]);
  } else {  // mean match emission when the file has no COMPO line
    for (int x = 0; x < K; x++) {
      double s = 0;
      for (int k = 1; k <= M; k++) s += mat[(size_t)k * K + x];
      compo[x] = (float)(s / M);
    }
  }
  cons.assign((size_t)M + 2, ' ');
  if (cons_) {
    for (int k = 1; k <= M; k++) cons[k] = cons_[k - 1];
  } else {
    float thr = abc == 0 ? 0.5f : 0.9f;
    for (int k = 1; k <= M; k++) {
      int best = 0;
      for (int x = 1; x < K; x++)
        if (mat[(size_t)k * K + x] > mat[(size_t)k * K + best]) best = x;
      char c = a.syms[best];
      cons[k] = mat[(size_t)k * K + best] >= thr ? c : (char)std::tolower(c);
    }
  }
  if (ev_) {
    std::memcpy(ev, ev_, sizeof(ev));
    has_stats = true;
  }
}

void HostProfile::configure() {
  const Alphabet &a = Alphabet::get(abc);
  const float NEG = -std::numeric_limits<float>::infinity();
  bgf.resize(K);
  for (int x = 0; x < K; x++) bgf[x] = abc == 0 ? kAminoBg[x] : 0.25f;
  tsc.assign((size_t)(M + 1) * 8, NEG);
  msc.assign((size_t)Kp * (M + 1), NEG);

  // local entry distribution from match-state occupancy
  std::vector<float> occ((size_t)M + 1, 0.0f);
  occ[1] = t[T_MI] + t[T_MM];
  for (int k = 2; k <= M; k++)
    occ[k] = (float)(occ[k - 1] * (t[(size_t)(k - 1) * 7 + T_MM] + t[(size_t)(k - 1) * 7 + T_MI]) +
                     (1.0 - occ[k - 1]) * t[(size_t)(k - 1) * 7 + T_DM]);
  float Z = 0.f;
  for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
  for (int k = 1; k <= M; k++) tsc[(size_t)(k - 1) * 8 + T_BM] = (float)std::log((double)(occ[k] / Z));
  for (int k = 1; k < M; k++)
    for (int x = 0; x < 7; x++) tsc[(size_t)k * 8 + x] = (float)std::log((double)t[(size_t)k * 7 + x]);
  if (M > 1) {
    tsc[8 + T_DM] = NEG;
    tsc[8 + T_DD] = NEG;
  }

  // match log-odds incl. degenerate residues (expected score), gap/'*'/'~' impossible
  std::vector<float> sc(Kp);
  for (int k = 1; k <= M; k++) {
    for (int x = 0; x < K; x++) sc[x] = (float)std::log((double)(mat[(size_t)k * K + x] / bgf[x]));
    sc[K] = sc[Kp - 2] = sc[Kp - 1] = NEG;
    for (int x = K + 1; x <= Kp - 3; x++) {
      uint32_t mem = a.members(x);
      float result = 0.f, denom = 0.f;
      for (int y = 0; y < K; y++)
        if (mem >> y & 1) {
          result += sc[y] * bgf[y];
          denom += bgf[y];
        }
      sc[x] = result / denom;
    }
    for (int x = 0; x < Kp; x++) msc[(size_t)x * (M + 1) + k] = sc[x];
  }

  // MSV bytes
  scale_b = (float)(3.0 / 0.69314718055994529);
  base_b = 190;
  float mx = 0.0f;
  for (int x = 0; x < K; x++)
    for (int k = 1; k <= M; k++) mx = std::fmax(mx, msc[(size_t)x * (M + 1) + k]);
  bias_b = unbiased_byteify(scale_b, -1.0f * mx);
  rbv.assign((size_t)Kp * (M + 1), 255);
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++)
      rbv[(size_t)x * (M + 1) + k] = biased_byteify(scale_b, bias_b, msc[(size_t)x * (M + 1) + k]);
  tbm_b = unbiased_byteify(scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
  tec_b = unbiased_byteify(scale_b, logf(0.5f));

  // Viterbi words
  scale_w = (float)(500.0 / 0.69314718055994529);
  base_w = 12000;
  rwv.assign((size_t)Kp * (M + 1), -32768);
  for (int x = 0; x < Kp; x++)
    for (int k = 1; k <= M; k++) rwv[(size_t)x * (M + 1) + k] = wordify(scale_w, msc[(size_t)x * (M + 1) + k]);
  twv.resize((size_t)(M + 1) * 8);
  for (int k = 0; k <= M; k++)
    for (int x = 0; x < 8; x++) {
      int16_t v = wordify(scale_w, tsc[(size_t)k * 8 + x]);
      int16_t cap = (x == T_II) ? -1 : 0;
      twv[(size_t)k * 8 + x] = v < cap ? v : cap;
    }
  xw_E = wordify(scale_w, -0.69314718055994529f);

  // float odds ratios
  rfv.resize(msc.size());
  tfv.resize(tsc.size());
  for (size_t i = 0; i < msc.size(); i++) rfv[i] = expf(msc[i]);
  for (size_t i = 0; i < tsc.size(); i++) tfv[i] = expf(tsc[i]);

  // bias filter state-1 odds
  for (int x = 0; x < 32; x++) eo1[x] = 1.0f;
  for (int x = 0; x < K; x++) eo1[x] = compo[x] / bgf[x];
  for (int x = K + 1; x <= Kp - 3; x++) {
    uint32_t mem = a.members(x);
    float num = 0.f, den = 0.f;
    for (int y = 0; y < K; y++)
      if (mem >> y & 1) {
        num += compo[y];
        den += bgf[y];
      }
    eo1[x] = num / den;
  }
}

double gumbel_surv(double x, double mu, double lambda) {
  double y = lambda * (x - mu);
  double ey = -std::exp(-y);
  if (std::fabs(ey) < 5e-9) return -ey;
  return 1.0 - std::exp(ey);
}
double exp_surv(double x, double mu, double lambda) {
  if (x < mu) return 1.0;
  return std::exp(-lambda * (x - mu));
}
double exp_logsurv(double x, double mu, double lambda) {
  if (x < mu) return 0.0;
  return -lambda * (x - mu);
}

static inline uint32_t fkey(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline float fromkey(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}
template <class Surv>
static float threshold_on_grid(double F, Surv surv) {
  const float INF = std::numeric_limits<float>::infinity();
  if (F >= 1.0) return -INF;  // P > F never true
  uint32_t lo = fkey(-3.0e38f), hi = fkey(3.0e38f);
  if (!(surv((double)fromkey(hi)) <= F)) return INF;  // nothing finite passes
  if (surv((double)fromkey(lo)) <= F) return fromkey(lo);
  while (hi - lo > 1) {  // invariant: surv(lo) > F, surv(hi) <= F
    uint32_t mid = lo + (hi - lo) / 2;
    if (surv((double)fromkey(mid)) <= F) hi = mid;
    else lo = mid;
  }
  return fromkey(hi);
}
float gumbel_threshold(double F, double mu, double lambda) {
  return threshold_on_grid(F, [&](double x) { return gumbel_surv(x, mu, lambda); });
}
float exp_threshold(double F, double mu, double lambda) {
  return threshold_on_grid(F, [&](double x) { return exp_surv(x, mu, lambda); });
}

float flogsum(float a, float b) {
  static float tbl[16000];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 16000; i++) tbl[i] = (float)std::log(1.0 + std::exp((double)-i / 1000.0));
    init = true;
  }
  float mx = a > b ? a : b, mn = a > b ? b : a;
  const float NEG = -std::numeric_limits<float>::infinity();
  return (mn == NEG || (mx - mn) >= 15.7f) ? mx : mx + tbl[(int)((mx - mn) * 1000.0f)];
}

}  // namespace pa

// ---------------------------------------------------------------------------------------------------------------------
// HMMER3/[a-f] ASCII reader (first model of a file): the numbers of `ref_hmm/{g}.hmm` as written by hmmbuild
// (/root/reference/lib/aln.py:66-76), SURVEY.md A.2.  Mirrors phyloaln_b200/hmmio.py::read_hmm_py field for field; the
// Python layer falls back to that statement for the descriptive error when this returns non-zero.
// ---------------------------------------------------------------------------------------------------------------------
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
struct Tok {  // whitespace tokenizer over one line
  const char *p, *e;
  bool next(const char *&b, size_t &n) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    if (p >= e) return false;
    b = p;
    while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
    n = (size_t)(p - b);
    return true;
  }
};
inline bool num(const char *b, size_t n, double *v) {
  if (n == 1 && *b == '*') { *v = INFINITY; return true; }
  {  // plain decimals (what hmmbuild writes): integer mantissa / exact power of ten is correctly rounded, i.e. == strtod
    static const double p10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    size_t z = 0;
    const bool neg = b[0] == '-';
    if (neg || b[0] == '+') z = 1;
    unsigned long long mant = 0;
    int digits = 0, frac = -1;
    bool ok = z < n;
    for (; z < n && ok; z++) {
      const char c = b[z];
      if (c >= '0' && c <= '9') { mant = mant * 10 + (unsigned)(c - '0'); digits++; if (frac >= 0) frac++; }
      else if (c == '.' && frac < 0) frac = 0;
      else ok = false;
    }
    if (ok && digits > 0 && digits <= 15 && frac <= 15) {
      const double x = (double)mant / p10[frac < 0 ? 0 : frac];
      *v = neg ? -x : x;
      return true;
    }
  }
  char buf[64];
  if (n >= sizeof(buf)) return false;
  memcpy(buf, b, n);
  buf[n] = 0;
  char *end = nullptr;
  *v = strtod(buf, &end);
  return end == buf + n;
}
}  // namespace

extern "C" int pa_hmm_file_parse(const char *path, int32_t *M_out, int32_t *abc_out, char *name, int namecap, double *mat,
                                 double *ins, double *t, double *compo, char *cons, char *rf, int32_t *map, float *stats,
                                 int32_t *nseq, double *effn) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return -1;
  std::string text;
  {
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, got);
  }
  fclose(fp);
  std::vector<std::pair<const char *, const char *>> lines;
  for (const char *p = text.data(), *e = p + text.size(); p < e;) {
    const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
    const char *le = nl ? nl : e;
    lines.emplace_back(p, le);
    p = nl ? nl + 1 : e;
  }
  if (lines.empty() || (size_t)(lines[0].second - lines[0].first) < 7 || memcmp(lines[0].first, "HMMER3/", 7) != 0) return -2;
  int M = -1, abc = -1, K = 0;
  bool has_cons = false, has_rf = false, has_map = false;
  float st[6];
  for (auto &x : st) x = NAN;
  if (name && namecap > 0) snprintf(name, (size_t)namecap, "unnamed");
  int32_t ns = 1;
  double ef = 1.0;
  size_t i = 1;
  for (; i < lines.size(); i++) {
    const char *b = lines[i].first, *e = lines[i].second;
    if (e - b >= 4 && memcmp(b, "HMM ", 4) == 0) break;
    Tok tk{b, e};
    const char *k0, *v0;
    size_t kn, vn;
    if (!tk.next(k0, kn) || !tk.next(v0, vn)) continue;
    const std::string key(k0, kn), val(v0, vn);
    if (key == "STATS") {
      const char *w[3];
      size_t wn[3];
      if (tk.next(w[0], wn[0]) && tk.next(w[1], wn[1]) && tk.next(w[2], wn[2])) {
        const std::string which(w[0], wn[0]);
        double a, c;
        if (num(w[1], wn[1], &a) && num(w[2], wn[2], &c)) {
          const int o = which == "MSV" ? 0 : which == "VITERBI" ? 2 : which == "FORWARD" ? 4 : -1;
          if (o >= 0) { st[o] = (float)a; st[o + 1] = (float)c; }
        } else return -3;
      }
    } else if (key == "LENG") M = atoi(val.c_str());
    else if (key == "ALPH") {
      std::string a = val;
      for (auto &c : a) c = (char)tolower((unsigned char)c);
      abc = a == "amino" ? 0 : (a == "dna" || a == "rna") ? 1 : -2;
    } else if (key == "NAME") { if (name && namecap > 0) snprintf(name, (size_t)namecap, "%s", val.c_str()); }
    else if (key == "CONS") has_cons = val == "yes" || val == "YES" || val == "Yes";
    else if (key == "RF") has_rf = val == "yes" || val == "YES" || val == "Yes";
    else if (key == "MAP") has_map = val == "yes" || val == "YES" || val == "Yes";
    else if (key == "NSEQ") ns = atoi(val.c_str());
    else if (key == "EFFN") ef = atof(val.c_str());
  }
  if (i >= lines.size() || M < 1 || abc < 0) return -3;
  K = abc == 0 ? 20 : 4;
  if (M_out) *M_out = M;
  if (abc_out) *abc_out = abc;
  if (nseq) *nseq = ns;
  if (effn) *effn = ef;
  if (stats) memcpy(stats, st, sizeof(st));
  if (!mat) return 0;   // size query
  i += 2;               // HMM line + transition-label line
  if (i >= lines.size()) return -3;
  auto numbers = [&](size_t li, int skip, int n, double *out) -> bool {
    if (li >= lines.size()) return false;
    Tok tk{lines[li].first, lines[li].second};
    const char *b;
    size_t bn;
    for (int s = 0; s < skip; s++)
      if (!tk.next(b, bn)) return false;
    for (int z = 0; z < n; z++)
      if (!tk.next(b, bn) || !num(b, bn, out + z)) return false;
    return true;
  };
  for (int z = 0; z < (M + 1) * K; z++) { mat[z] = INFINITY; ins[z] = INFINITY; }
  for (int z = 0; z < (M + 1) * 7; z++) t[z] = INFINITY;
  for (int z = 0; z < K; z++) compo[z] = NAN;
  {
    Tok tk{lines[i].first, lines[i].second};
    const char *b;
    size_t bn;
    if (tk.next(b, bn) && bn == 5 && memcmp(b, "COMPO", 5) == 0) {
      if (!numbers(i, 1, K, compo)) return -3;
      i++;
    }
  }
  if (!numbers(i, 0, K, ins) || !numbers(i + 1, 0, 7, t)) return -3;
  i += 2;
  cons[0] = rf[0] = 0;
  int ncons = 0, nrf = 0, nmap = 0;
  for (int k = 1; k <= M; k++, i += 3) {
    if (i + 2 >= lines.size()) return -3;
    Tok tk{lines[i].first, lines[i].second};
    const char *b;
    size_t bn;
    if (!tk.next(b, bn)) return -3;
    char kb[16];
    const int kl = snprintf(kb, sizeof(kb), "%d", k);
    if ((size_t)kl != bn || memcmp(kb, b, bn) != 0) return -3;
    for (int z = 0; z < K; z++)
      if (!tk.next(b, bn) || !num(b, bn, &mat[(size_t)k * K + z])) return -3;
    if (tk.next(b, bn)) {                       // MAP
      if (has_map) { map[nmap++] = (bn == 1 && *b == '-') ? 0 : atoi(std::string(b, bn).c_str()); }
      if (tk.next(b, bn)) {                     // CONS
        if (has_cons) { cons[ncons++] = *b; }
        if (tk.next(b, bn) && has_rf) rf[nrf++] = *b;
      }
    }
    if (!numbers(i + 1, 0, K, ins + (size_t)k * K) || !numbers(i + 2, 0, 7, t + (size_t)k * 7)) return -3;
  }
  cons[ncons == M ? M : 0] = 0;
  rf[nrf == M ? M : 0] = 0;
  if (nmap != M) map[0] = -1;   // absent
  return 0;
}
