// api.cu -- C ABI of libphyloaln_b200.so (see include/phyloaln_b200.h) and the host-side
// orchestration of the kernels in kernels.cuh / domain.cuh / mapping.cuh.
//
// Flow of pa_search for one resident chunk of targets and ALL committed profiles (the reference
// instead runs one hmmsearch process per profile over all targets, lib/aln.py:664-667):
//   msv_kernel<Q> (every pair)  -> Pass1 list  -> bias_kernel -> Pass2 list -> vit_kernel ->
//   Pass2 list -> fwd_kernel -> Pass4 list -> domain_kernel -> domain records -> host scoring.
// Lists are compacted with atomics between stages; capacity overflow is reported as an error
// (-3) so the caller can shrink the chunk -- there is no CPU fallback anywhere.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phyloaln_b200.h"
#include "kernels.cuh"
#include "msv_tmem.cuh"
#include "vit16.cuh"
#include "wide.cuh"
#include "launch.hpp"
#include "domain.cuh"
#include "gencode.hpp"
#include "ctx_access.hpp"
#include "profile.hpp"

using namespace pa;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof(b_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = b_;                                                                               \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct pa_ctx {
  int device = 0;
  std::string err;
  pa_opts opts{0.02, 1e-3, 1e-5, 1, 1, 0, 0};
  cudaStream_t stream = nullptr;
  cudaStream_t side[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // concurrent per-class launches
  cudaEvent_t side_ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int num_sms = 148;
  size_t smem_optin = 0;

  std::vector<HostProfile> hmms;
  bool committed = false;
  bool tables_shared = false;  // profiles came from pa_commit_hmms_multi: only a slim copy lives here, no re-commit
  std::vector<DevHmm> h_dev;
  DevBuf<DevHmm> d_hmms;
  DevBuf<uint32_t> d_msv, d_msvh;
  int msv_impl = 0;  // 0: TMEM kernel where the table fits (default), 1: shared-memory kernel everywhere (PA_MSV_IMPL=smem)
  bool ftok = false; // the alphabet's table fits TMEM at 26 words per row (DNA): all-TMEM rows for Q 17..26, 1,664-cell packs, all-TMEM wide tiles
  int pack_q = 16;   // words per lane of a pack (16, or 26 with ftok)
  DevBuf<int> d_vit;
  DevBuf<uint32_t> d_v16;   // packed tables of vit16_kernel
  DevBuf<uint8_t> d_need;   // per sorted Viterbi input: must the exact kernel run?
  struct VClass { int B, NW, r0, r1; };   // block size (class-rounded), warps per pair (1: single-warp kernels), rank range [r0, r1)
  std::vector<VClass> v16_classes;
  int use_vit16 = 1;        // PA_VIT16=0: exact kernel on every pair
  DevBuf<float> d_fwd;
  DevBuf<int> d_hmm_lists;
  std::map<int, std::pair<int, int>> q_groups;  // Q -> (offset, count) in d_hmm_lists (all profiles: score_all / single-profile runs)
  std::map<int, std::pair<int, int>> q_groups_rest;  // same, profiles that are not in a pack (d_hmm_lists_rest)
  DevBuf<int> d_hmm_lists_rest;
  std::map<int, std::pair<int, int>> w_groups;  // wide profiles (M > 2047): tile width QW -> (offset, count) in d_hmm_lists_w
  DevBuf<int> d_hmm_lists_w;
  int maxQ_wide = 0;         // widest striped state (words per lane) of msv_exact_wide_kernel
  DevBuf<uint16_t> d_bnd;    // tile-boundary scratch of msv_wide_kernel
  DevBuf<Pass1> d_wcand;     // its candidate list
  DevBuf<float> d_fwdw;      // per-CTA P rows of fwd_wide_kernel
  std::vector<DevPack> h_packs;   // packed narrow profiles (msv_pack_kernel)
  DevBuf<DevPack> d_packs;
  DevBuf<int> d_packmem;
  int maxB = 1, maxM = 1;
  std::vector<VClass> vit_classes;
  std::vector<int> rank_B;  // canonical Forward block size B of the profile at each rank (ranks are sorted by B)
  DevBuf<unsigned int> d_sort;  // cnt | start | cursor, each nh + 1

  // targets
  DevBuf<char> d_nt;
  DevBuf<unsigned long long> d_ntoff;
  DevBuf<uint8_t> d_dsq;
  DevBuf<uint32_t> d_toff, d_tlen;
  DevBuf<uint16_t> d_lidx;
  DevBuf<uint8_t> d_codon;
  DevBuf<int> d_flag;
  std::vector<uint32_t> h_tlen;
  uint32_t ntargets = 0;
  int nframes = 1;
  int target_abc = 0;
  int maxL = 0;
  unsigned long long total_res = 0;

  // per-length tables
  std::vector<int> h_L;
  DevBuf<int> d_L;
  DevBuf<float> d_nullsc, d_biasA, d_biasB;
  DevBuf<uint8_t> d_tjb;
  DevBuf<int16_t> d_xwmove, d_thrJ;
  DevBuf<uint16_t> d_scJ;
  DevBuf<uint16_t> d_L2li;
  int nL = 0;

  // stage lists
  DevBuf<Pass1> d_p1;
  DevBuf<Pass2> d_p2a, d_p2b;
  DevBuf<Pass4> d_p4;
  DevBuf<unsigned long long> d_msvwork;   // per MSV launch: next work item
  DevBuf<unsigned long long> d_counters;  // [0]=p1 [1]=p2a [2]=p2b [3]=p4 [4]=ssv cand [5..] domain
  DomainBufs dom;

  std::vector<pa_hit> hits;
  std::string model_pool, target_pool, pp_pool;
  pa_stats stats{};
};

static Targets make_targets(pa_ctx *ctx) {
  Targets t;
  t.dsq = ctx->d_dsq.p;
  t.off = ctx->d_toff.p;
  t.len = ctx->d_tlen.p;
  t.lidx = ctx->d_lidx.p;
  t.n = ctx->ntargets;
  return t;
}
static LenTabs make_lentabs(pa_ctx *ctx) {
  LenTabs l;
  l.L = ctx->d_L.p;
  l.nullsc = ctx->d_nullsc.p;
  l.tjb = ctx->d_tjb.p;
  l.xw_move = ctx->d_xwmove.p;
  l.biasA = ctx->d_biasA.p;
  l.biasB = ctx->d_biasB.p;
  l.n = ctx->nL;
  return l;
}

extern "C" int pa_abi_version(void) { return PA_ABI_VERSION; }

extern "C" int pa_profiler_range(int on) { return (on ? cudaProfilerStart() : cudaProfilerStop()) == cudaSuccess ? 0 : -1; }

extern "C" pa_ctx *pa_create(int device, char *err, int errlen) {
  auto fail = [&](const char *what, cudaError_t e) -> pa_ctx * {
    if (err && errlen > 0) snprintf(err, errlen, "%s: %s", what, cudaGetErrorString(e));
    return nullptr;
  };
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("no CUDA device (this library has no CPU fallback)", e);
  if (device < 0 || device >= ndev) return fail("bad device index", cudaErrorInvalidDevice);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (prop.major < 9) {
    if (err && errlen > 0) snprintf(err, errlen, "device %d is sm_%d%d; DPX kernels need sm_100a (B200)", device, prop.major, prop.minor);
    return nullptr;
  }
  pa_ctx *ctx = new pa_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (const char *impl = getenv("PA_MSV_IMPL")) ctx->msv_impl = !strcmp(impl, "smem") ? 1 : 0;
  if (const char *v = getenv("PA_VIT16")) ctx->use_vit16 = strcmp(v, "0") != 0;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return fail("cudaStreamCreate", e);
  }
  return ctx;
}

extern "C" void pa_destroy(pa_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  ctx->d_hmms.release(); ctx->d_msv.release(); ctx->d_msvh.release(); ctx->d_vit.release(); ctx->d_v16.release(); ctx->d_need.release(); ctx->d_fwd.release(); ctx->d_hmm_lists.release(); ctx->d_hmm_lists_rest.release(); ctx->d_hmm_lists_w.release(); ctx->d_bnd.release(); ctx->d_wcand.release(); ctx->d_fwdw.release(); ctx->d_packs.release(); ctx->d_packmem.release(); ctx->d_sort.release();
  ctx->d_nt.release(); ctx->d_ntoff.release(); ctx->d_dsq.release(); ctx->d_toff.release(); ctx->d_tlen.release();
  ctx->d_lidx.release(); ctx->d_codon.release(); ctx->d_flag.release();
  ctx->d_L.release(); ctx->d_nullsc.release(); ctx->d_biasA.release(); ctx->d_biasB.release(); ctx->d_tjb.release();
  ctx->d_xwmove.release(); ctx->d_thrJ.release(); ctx->d_scJ.release(); ctx->d_L2li.release();
  ctx->d_p1.release(); ctx->d_p2a.release(); ctx->d_p2b.release(); ctx->d_p4.release(); ctx->d_counters.release(); ctx->d_msvwork.release();
  ctx->dom.release();
  for (auto &st : ctx->side) if (st) cudaStreamDestroy(st);
  for (auto &ev : ctx->side_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *pa_last_error(pa_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int pa_set_opts(pa_ctx *ctx, const pa_opts *o) {
  if (!ctx || !o) return -1;
  ctx->opts = *o;
  if (ctx->committed) {  // thresholds live in DevHmm: refresh
    ctx->committed = false;
    return pa_commit_hmms(ctx);
  }
  return 0;
}

extern "C" int pa_add_hmm(pa_ctx *ctx, const char *name, int abc, int M, const double *mat, const double *ins,
                          const double *t, const double *compo, const char *cons, const float *evparam) {
  if (!ctx) return -1;
  if (M < 1 || M > PA_MAX_M) { ctx->err = "profile length must be 1.." + std::to_string(PA_MAX_M); return -2; }
  if (abc != PA_AMINO && abc != PA_DNA) { ctx->err = "unknown alphabet"; return -2; }
  if (!ctx->hmms.empty() && ctx->hmms[0].abc != abc) { ctx->err = "all profiles of a context must share one alphabet"; return -2; }
  if (ctx->hmms.size() >= 65535) { ctx->err = "too many profiles (max 65535)"; return -2; }
  HostProfile hp;
  hp.from_file_numbers(name, abc, M, mat, ins, t, compo, cons, evparam);   // configured at commit time, all profiles in parallel
  ctx->hmms.push_back(std::move(hp));
  ctx->committed = false;
  return (int)ctx->hmms.size() - 1;
}

extern "C" int pa_set_evparam(pa_ctx *ctx, int h, const float *ev) {
  if (!ctx || h < 0 || h >= (int)ctx->hmms.size()) return -1;
  memcpy(ctx->hmms[h].ev, ev, sizeof(float) * 6);
  ctx->hmms[h].has_stats = true;
  ctx->committed = false;
  return 0;
}

extern "C" int pa_num_hmms(pa_ctx *ctx) { return ctx ? (int)ctx->hmms.size() : -1; }

static int v16_class(int B) {
  static const int cls[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32};
  for (int c : cls)
    if (B <= c) return c;
  return 32;
}
static inline int clamp_v16(int v) { return v < PA_V16_FLOOR ? PA_V16_FLOOR : v; }

static int vit_class(int B) {
  static const int cls[] = {1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64};
  for (int c : cls)
    if (B <= c) return c;
  return 64;
}

// word position of (q, lane) inside one striped MSV row of Q*32 words (see msv_load_row)
static inline int msv_word_pos(int Q, int q, int lane) {
  const int N4 = Q / 4, R2 = (Q % 4) >= 2 ? 1 : 0;
  if (q < 4 * N4) return (q / 4) * 128 + lane * 4 + (q % 4);
  if (R2 && q < 4 * N4 + 2) return N4 * 128 + lane * 2 + (q - 4 * N4);
  return N4 * 128 + R2 * 64 + lane;
}

// exact fp16 bit pattern of an integer that fp16 represents exactly (|v| <= 2048, or a multiple of 16 up to 32768)
static inline uint16_t h16_exact(int v) {
  if (v == 0) return 0;
  const float f = (float)v;
  uint32_t b;
  memcpy(&b, &f, 4);
  const uint32_t sign = b >> 31, e = ((b >> 23) & 0xff) - 127 + 15, mant = (b >> 13) & 0x3ff;
  return (uint16_t)((sign << 15) | (e << 10) | mant);
}

static inline bool msv_fullt(const pa_ctx *ctx, int Q) { return ctx->ftok && Q > 16 && Q <= PA_TMEM_FULLQ; }
static inline bool msv_use_tmem(const pa_ctx *ctx, int Q, int nrows) {
  return ctx->msv_impl == 0 && Q <= PA_TMEM_MAXQ && msv_tmem_cols_needed(Q, nrows, msv_fullt(ctx, Q)) <= 512;
}

// fn(i) for i in [0, n) on the host's hardware threads (table building is independent per profile)
template <class F>
static void parallel_for(int n, F fn) {
  const int nt = std::max(1, std::min<int>({n, (int)std::thread::hardware_concurrency(), 16}));
  if (nt <= 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}

// Host images of every per-profile device table (built once per profile set, uploaded to one or several GPUs)
struct HostTables {
  std::vector<uint32_t> msv, msvh, v16;
  std::vector<int> vit, packmem, lists_rest, lists_all, lists_w;
  std::vector<float> fwd;
};

static int build_tables(pa_ctx *ctx, HostTables &T) {
  const int nh = (int)ctx->hmms.size();
  if (nh == 0) { ctx->err = "no profiles"; return -2; }
  if (ctx->tables_shared) { ctx->err = "this context shares the profiles of another one (pa_commit_hmms_multi): change options or profiles on the first context and commit again"; return -2; }
  std::vector<uint32_t> &msv = T.msv, &msvh = T.msvh, &v16 = T.v16;
  std::vector<int> &vit = T.vit;
  std::vector<float> &fwd = T.fwd;
  ctx->h_dev.assign(nh, DevHmm{});
  ctx->maxB = 1;
  ctx->maxM = 1;
  std::map<int, std::vector<int>> byQ, byW;
  ctx->maxQ_wide = 0;
  // All-TMEM rows of up to 26 words for small alphabets are implemented and parity-tested, but OFF by default: measured on
  // config 3 (DNA) they are 6 % slower than the hybrid rows (8,960 vs 8,462 ms of MSV per 500 k reads) -- one LDTM.x32 per
  // 1,664-cell row is 4 KB, i.e. ~315 B/clk/SM at the FMA-bound rate, which is the measured TMEM read ceiling (270-400 B/clk/SM,
  // tools/probe_sm100.cu), and the 128-register budget leaves a single row of prefetch.  PA_MSV_FULLT=1 switches them on.
  ctx->ftok = ctx->msv_impl == 0 && getenv("PA_MSV_FULLT") != nullptr && atoi(getenv("PA_MSV_FULLT")) != 0 &&
              msv_tmem_full_ok(PA_TMEM_FULLQ, ctx->hmms[0].Kp);
  ctx->pack_q = ctx->ftok ? PA_TMEM_FULLQ : 16;
  const bool maxmode = ctx->opts.do_max != 0;
  parallel_for(nh, [&](int h) { if (!ctx->hmms[h].configured) { ctx->hmms[h].configure(); ctx->hmms[h].configured = true; } });
  size_t n_msv = 0, n_msvh = 0, n_vit = 0, n_v16 = 0, n_fwd = 0;
  std::vector<int> qw_of(nh, 0);
  for (int h = 0; h < nh; h++) {
    const HostProfile &hp = ctx->hmms[h];
    if (!hp.has_stats) { ctx->err = "profile '" + hp.name + "' has no STATS LOCAL (uncalibrated)"; return -2; }
    DevHmm &d = ctx->h_dev[h];
    const int M = hp.M, Kp = hp.Kp;
    d.M = M;
    d.abc = hp.abc;
    d.nrows = Kp;
    d.Q = M / 64 + 1;  // the last striped cell must be padding (see msv_kernel)
    d.B = (M + 31) / 32;
    const bool wide = d.Q > 32;  // wide.cuh: model-tiled MSV, multi-warp Viterbi, run-time-B Forward
    const int tileQ = ctx->ftok ? PA_TMEM_FULLQ : 32;   // widest tile: all-TMEM rows for small alphabets, hybrid otherwise
    d.nT = wide ? (d.Q + tileQ - 1) / tileQ : 0;
    const int QW = wide ? (d.Q + d.nT - 1) / d.nT : 0;  // words per lane of one MSV tile, 17..32 (17..26 all-TMEM)
    d.NWv = wide ? (M <= 4096 ? 4 : 8) : 1;
    d.NW16 = wide ? (M <= 4096 ? 2 : 4) : 1;
    ctx->maxB = std::max(ctx->maxB, d.B);
    ctx->maxM = std::max(ctx->maxM, M);
    if (wide) { byW[QW].push_back(h); ctx->maxQ_wide = std::max(ctx->maxQ_wide, d.Q); }
    else byQ[d.Q].push_back(h);
    d.tbm = hp.tbm_b; d.tec = hp.tec_b; d.bias = hp.bias_b; d.base = hp.base_b;
    d.base_w = hp.base_w; d.xw_E = hp.xw_E; d.scale_b = hp.scale_b; d.scale_w = hp.scale_w;
    const double F1 = maxmode ? 1.0 : ctx->opts.F1, F2 = maxmode ? 1.0 : ctx->opts.F2, F3 = maxmode ? 1.0 : ctx->opts.F3;
    d.thrF1 = gumbel_threshold(F1, hp.ev[0], hp.ev[1]);
    d.thrF2m = gumbel_threshold(F2, hp.ev[0], hp.ev[1]);
    d.thrF2v = gumbel_threshold(F2, hp.ev[2], hp.ev[3]);
    d.thrF3 = exp_threshold(F3, hp.ev[4], hp.ev[5]);
    const float L0 = 400.0f, L1 = (float)M / 8.0f;
    d.t00 = L0 / (L0 + 1.0f); d.t01 = 1.0f / (L0 + 1.0f);
    d.t10 = 1.0f / (L1 + 1.0f); d.t11 = L1 / (L1 + 1.0f);
    memcpy(d.eo1, hp.eo1, sizeof(d.eo1));
    memcpy(d.ev, hp.ev, sizeof(d.ev));

    // ---- sizes and offsets of this profile's tables (filled in parallel below)
    auto take = [](size_t &total, size_t n) { total = (total + 3) & ~(size_t)3; const size_t o = total; total += n; return o; };
    d.msv_off = take(n_msv, (size_t)Kp * d.Q * 32);
    d.msvh_off = n_msvh;
    if (wide) {
      if (msv_tmem_cols_needed(QW, Kp, ctx->ftok) > 512) { ctx->err = "alphabet too large for the tiled MSV kernel"; return -2; }
      n_msvh += (size_t)Kp * QW * 32 * d.nT;
    } else if (msv_use_tmem(ctx, d.Q, Kp))
      n_msvh += (size_t)Kp * d.Q * 32;
    d.Bv = vit_class((M + 32 * d.NWv - 1) / (32 * d.NWv));
    d.vit_off = take(n_vit, (size_t)d.Bv * 32 * d.NWv * (9 + Kp));
    d.BH = v16_class((M + 64 * d.NW16 - 1) / (64 * d.NW16));
    d.v16_off = take(n_v16, (size_t)d.BH * 32 * d.NW16 * (9 + Kp));
    d.fwd_off = take(n_fwd, (size_t)d.B * 32 * (16 + Kp));
    qw_of[h] = QW;
  }
  msv.assign(n_msv, 0u);
  msvh.assign(n_msvh, 0u);
  vit.assign(n_vit, PA_NEG16);
  v16.assign(n_v16, 0u);
  fwd.assign(n_fwd, 0.0f);
  auto fill_profile = [&](int h) {
    const HostProfile &hp = ctx->hmms[h];
    DevHmm &d = ctx->h_dev[h];
    const int M = hp.M, Kp = hp.Kp;
    const bool wide = d.nT > 0;
    const int QW = qw_of[h];
    // ---- MSV striped s16x2 table: delta = bias - rbv, dead padding beyond M (wide profiles: plain [x][q][lane] rows)
    const int Q = d.Q, roww = Q * 32;
    for (int x = 0; x < Kp; x++)
      for (int q = 0; q < Q; q++)
        for (int lane = 0; lane < 32; lane++) {
          int16_t half[2];
          for (int hf = 0; hf < 2; hf++) {
            const int cell = (2 * lane + hf) * Q + q, k = cell + 1;
            half[hf] = k <= M ? (int16_t)((int)hp.bias_b - (int)hp.rbv[(size_t)x * (M + 1) + k]) : (int16_t)PA_MSV_DEAD;
          }
          msv[d.msv_off + (size_t)x * roww + (wide ? q * 32 + lane : msv_word_pos(Q, q, lane))] = (uint32_t)(uint16_t)half[0] | ((uint32_t)(uint16_t)half[1] << 16);
        }

    // ---- the same deltas as fp16x2 for the tensor-memory kernel (msv_tmem.cuh): words q < QT as
    // [x][q][lane] (TMEM part), words q >= QT striped like the s16 table (shared-memory part, Q > 16)
    if (wide) {
      // nT tiles of 64*QW cells, each laid out like a profile with Q = QW: tile m holds model cells m*64*QW + (2*lane + hf)*QW + q
      const int QT = msv_tmem_qt(QW, ctx->ftok), QS = QW - QT;
      const size_t tilew = (size_t)Kp * QW * 32;
      for (int m = 0; m < d.nT; m++) {
        uint32_t *tpart = msvh.data() + d.msvh_off + tilew * m, *spart = tpart + (size_t)Kp * QT * 32;
        for (int x = 0; x < Kp; x++)
          for (int q = 0; q < QW; q++)
            for (int lane = 0; lane < 32; lane++) {
              uint16_t half[2];
              for (int hf = 0; hf < 2; hf++) {
                const int cell = m * 64 * QW + (2 * lane + hf) * QW + q, k = cell + 1;
                half[hf] = h16_exact(k <= M ? (int)hp.bias_b - (int)hp.rbv[(size_t)x * (M + 1) + k] : PA_MSV_DEAD);
              }
              const uint32_t word = (uint32_t)half[0] | ((uint32_t)half[1] << 16);
              if (q < QT) tpart[((size_t)x * QT + q) * 32 + lane] = word;
              else spart[(size_t)x * QS * 32 + msv_word_pos(QS, q - QT, lane)] = word;
            }
      }
    } else if (msv_use_tmem(ctx, Q, Kp)) {
      const int QT = msv_tmem_qt(Q, msv_fullt(ctx, Q)), QS = Q - QT;
      uint32_t *tpart = msvh.data() + d.msvh_off, *spart = tpart + (size_t)Kp * QT * 32;
      for (int x = 0; x < Kp; x++)
        for (int q = 0; q < Q; q++)
          for (int lane = 0; lane < 32; lane++) {
            uint16_t half[2];
            for (int hf = 0; hf < 2; hf++) {
              const int cell = (2 * lane + hf) * Q + q, k = cell + 1;
              half[hf] = h16_exact(k <= M ? (int)hp.bias_b - (int)hp.rbv[(size_t)x * (M + 1) + k] : PA_MSV_DEAD);
            }
            const uint32_t word = (uint32_t)half[0] | ((uint32_t)half[1] << 16);
            if (q < QT) tpart[((size_t)x * QT + q) * 32 + lane] = word;
            else spart[(size_t)x * QS * 32 + msv_word_pos(QS, q - QT, lane)] = word;
          }
    }

    // ---- Viterbi tables, thread-blocked [j][tid] with the class-rounded block size Bv: k = tid*Bv + j + 1
    // (tid < NT: one warp, or the NWv warps of a CTA for wide profiles)
    {
      const int NT = 32 * d.NWv;
      const int Bv = d.Bv, Wv = Bv * NT;
      int *vA = vit.data() + d.vit_off, *vI = vA + 4 * Wv, *vD = vI + 2 * Wv, *vps = vD + 2 * Wv, *vem = vps + Wv;
      for (int c = 0; c < Wv; c++) {
        const int k = (c % NT) * Bv + (c / NT) + 1;
        if (k <= M) {
          const int16_t *tp = &hp.twv[(size_t)(k - 1) * 8], *tk = &hp.twv[(size_t)k * 8];
          vA[4 * c + 0] = tp[T_BM]; vA[4 * c + 1] = tp[T_MM]; vA[4 * c + 2] = tp[T_IM]; vA[4 * c + 3] = tp[T_DM];
          vI[2 * c + 0] = tk[T_MI]; vI[2 * c + 1] = tk[T_II];
          vD[2 * c + 0] = tp[T_MD]; vD[2 * c + 1] = tp[T_DD];
          for (int x = 0; x < Kp; x++) vem[(size_t)x * Wv + c] = hp.rwv[(size_t)x * (M + 1) + k];
        }
      }
      for (int tid = 0; tid < NT; tid++) {  // in-thread prefix sums of DD(k-1)
        int sum = 0;
        for (int j = 0; j < Bv; j++) {
          const int c = j * NT + tid;
          sum += vD[2 * c + 1];
          vps[c] = sum;
        }
      }
    }
    // ---- packed tables of the 16-bit upper-bound pass: half-lane hl = 2*tid + hf owns k = hl*BH + j + 1; every entry >= -16384
    {
      const int NT = 32 * d.NW16;
      const int BH = d.BH, W16 = BH * NT;
      uint32_t *wA = v16.data() + d.v16_off, *wI = wA + 4 * W16, *wD = wI + 2 * W16, *wps = wD + 2 * W16, *wem = wps + W16;
      int guard = 0;
      auto put = [](uint32_t *w, int hf, int v) { *w = hf ? ((*w & 0x0000ffffu) | ((uint32_t)(v & 0xffff) << 16)) : ((*w & 0xffff0000u) | (uint32_t)(v & 0xffff)); };
      for (int c = 0; c < W16; c++)
        for (int hf = 0; hf < 2; hf++) {
          const int k = (2 * (c % NT) + hf) * BH + (c / NT) + 1;
          int tA[4] = {PA_V16_FLOOR, PA_V16_FLOOR, PA_V16_FLOOR, PA_V16_FLOOR}, tI[2] = {PA_V16_FLOOR, PA_V16_FLOOR},
              tD[2] = {PA_V16_FLOOR, PA_V16_FLOOR};
          if (k <= M) {
            const int16_t *tp = &hp.twv[(size_t)(k - 1) * 8], *tk = &hp.twv[(size_t)k * 8];
            tA[0] = clamp_v16(tp[T_BM]); tA[1] = clamp_v16(tp[T_MM]); tA[2] = clamp_v16(tp[T_IM]); tA[3] = clamp_v16(tp[T_DM]);
            tI[0] = clamp_v16(tk[T_MI]); tI[1] = clamp_v16(tk[T_II]);
            tD[0] = clamp_v16(tp[T_MD]); tD[1] = clamp_v16(tp[T_DD]);
          }
          for (int z = 0; z < 4; z++) put(&wA[4 * c + z], hf, tA[z]);
          for (int z = 0; z < 2; z++) { put(&wI[2 * c + z], hf, tI[z]); put(&wD[2 * c + z], hf, tD[z]); }
          for (int x = 0; x < Kp; x++) {
            const int e = k <= M ? clamp_v16(hp.rwv[(size_t)x * (M + 1) + k]) : PA_V16_FLOOR;
            put(&wem[(size_t)x * W16 + c], hf, e);
            guard = std::max(guard, e);
          }
        }
      for (int tid = 0; tid < NT; tid++)
        for (int hf = 0; hf < 2; hf++) {  // in-half-lane prefix sums of DD(k-1), floored
          int sum = 0;
          for (int j = 0; j < BH; j++) {
            const int c = j * NT + tid;
            sum = clamp_v16(sum + (int)(short)(hf ? (wD[2 * c + 1] >> 16) : (wD[2 * c + 1] & 0xffff)));
            put(&wps[c], hf, sum);
          }
        }
      d.v16_guard = guard + std::abs((int)d.xw_E) + 8;
    }
    // ---- float tables, lane-blocked [j][lane] with the canonical B = ceil(M/32): k = lane*B + j + 1
    const int B = d.B, W = B * 32;
    auto cellk = [&](int c) { return (c % 32) * B + (c / 32) + 1; };
    float *fA = fwd.data() + d.fwd_off, *fB = fA + 4 * W, *bA = fB + 4 * W, *bB = bA + 4 * W, *fem = bB + 4 * W;
    for (int c = 0; c < W; c++) {
      const int k = cellk(c);
      if (k <= M) {
        const float *tp = &hp.tfv[(size_t)(k - 1) * 8], *tk = &hp.tfv[(size_t)k * 8];
        fA[4 * c + 0] = tp[T_BM]; fA[4 * c + 1] = tp[T_MM]; fA[4 * c + 2] = tp[T_IM]; fA[4 * c + 3] = tp[T_DM];
        fB[4 * c + 0] = tp[T_MD]; fB[4 * c + 1] = tp[T_DD]; fB[4 * c + 2] = tk[T_MI]; fB[4 * c + 3] = tk[T_II];
        bA[4 * c + 0] = tk[T_MM]; bA[4 * c + 1] = tk[T_MI]; bA[4 * c + 2] = tk[T_MD]; bA[4 * c + 3] = tk[T_IM];
        bB[4 * c + 0] = tk[T_II]; bB[4 * c + 1] = tk[T_DM]; bB[4 * c + 2] = tk[T_DD]; bB[4 * c + 3] = tp[T_BM];
        for (int x = 0; x < Kp; x++) fem[(size_t)x * W + c] = hp.rfv[(size_t)x * (M + 1) + k];
      }
    }
  };
  parallel_for(nh, fill_profile);
  {  // profile rank in (M, index) order: the sorted survivor list is then contiguous per Viterbi class and per Forward block size
    std::vector<int> ord(nh);
    for (int h = 0; h < nh; h++) ord[h] = h;
    // by M: B = ceil(M/32), the Viterbi classes (block size and warps per pair) and the vit16 classes are all monotone in M
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return ctx->h_dev[x].M < ctx->h_dev[y].M; });
    ctx->rank_B.assign(nh, 1);
    ctx->vit_classes.clear();
    ctx->v16_classes.clear();
    for (int r = 0; r < nh; r++) {
      DevHmm &d = ctx->h_dev[ord[r]];
      d.rank = r;
      ctx->rank_B[r] = d.B;
      if (ctx->vit_classes.empty() || ctx->vit_classes.back().B != d.Bv || ctx->vit_classes.back().NW != d.NWv) ctx->vit_classes.push_back({d.Bv, d.NWv, r, r + 1});
      else ctx->vit_classes.back().r1 = r + 1;
      if (ctx->v16_classes.empty() || ctx->v16_classes.back().B != d.BH || ctx->v16_classes.back().NW != d.NW16) ctx->v16_classes.push_back({d.BH, d.NW16, r, r + 1});
      else ctx->v16_classes.back().r1 = r + 1;
    }
  }
  // ---- packs of narrow profiles for msv_pack_kernel (first-fit decreasing into 64 half-lanes of 16 cells)
  std::vector<int> &packmem = T.packmem;
  std::vector<char> packed(nh, 0);
  ctx->h_packs.clear();
  if (ctx->msv_impl == 0 && getenv("PA_MSV_NOPACK") == nullptr) {
    const int Kp0 = ctx->hmms[0].Kp;
    const int QP = ctx->pack_q;   // cells per half-lane = words per lane of the packed model
    if (msv_tmem_cols_needed(QP, Kp0, ctx->ftok) <= 512) {
      std::vector<int> cand;
      for (int h = 0; h < nh; h++)
        if (ctx->hmms[h].M / QP + 1 <= 64) cand.push_back(h);
      std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) { return ctx->hmms[x].M > ctx->hmms[y].M; });
      std::vector<std::vector<int>> bins;
      std::vector<int> fill;
      for (int h : cand) {
        const int n = ctx->hmms[h].M / QP + 1;
        size_t b = 0;
        while (b < bins.size() && fill[b] + n > 64) b++;
        if (b == bins.size()) { bins.emplace_back(); fill.push_back(0); }
        bins[b].push_back(h);
        fill[b] += n;
      }
      for (auto &bin : bins) {
        DevPack pk;
        pk.nrows = Kp0;
        pk.nmem = (int)bin.size();
        pk.mem_off = (int)packmem.size();
        while (msvh.size() % 4) msvh.push_back(0);
        pk.msvh_off = msvh.size();
        msvh.resize(msvh.size() + (size_t)Kp0 * QP * 32, 0);
        uint32_t *tab = msvh.data() + pk.msvh_off;
        std::vector<int> owner(64, -1), first(64, 0);
        int s = 0;
        for (int h : bin) {
          const int n = ctx->hmms[h].M / QP + 1;
          packmem.push_back(h); packmem.push_back(s); packmem.push_back(n);
          for (int z = 0; z < n; z++) { owner[s + z] = h; first[s + z] = s; }
          s += n;
          packed[h] = 1;
        }
        for (int hl = 0; hl < 64; hl++) packmem.push_back(owner[hl]);  // per half-lane owner, read by msv_pack_kernel
        for (int x = 0; x < Kp0; x++)
          for (int q = 0; q < QP; q++)
            for (int lane = 0; lane < 32; lane++) {
              uint16_t half[2];
              for (int hf = 0; hf < 2; hf++) {
                const int hl = 2 * lane + hf, h = owner[hl];
                int delta = PA_MSV_DEAD;
                if (h >= 0) {
                  const HostProfile &hp = ctx->hmms[h];
                  const int k = (hl - first[hl]) * QP + q + 1;
                  if (k <= hp.M) delta = (int)hp.bias_b - (int)hp.rbv[(size_t)x * (hp.M + 1) + k];
                }
                half[hf] = h16_exact(delta);
              }
              tab[((size_t)x * QP + q) * 32 + lane] = (uint32_t)half[0] | ((uint32_t)half[1] << 16);
            }
        ctx->h_packs.push_back(pk);
      }
    }
  }
  {  // per-Q lists of the profiles that stay on the per-profile kernels
    std::map<int, std::vector<int>> rest;
    for (int h = 0; h < nh; h++)
      if (!packed[h] && ctx->h_dev[h].nT == 0) rest[ctx->h_dev[h].Q].push_back(h);
    ctx->q_groups_rest.clear();
    for (auto &kv : rest) {
      ctx->q_groups_rest[kv.first] = {(int)T.lists_rest.size(), (int)kv.second.size()};
      T.lists_rest.insert(T.lists_rest.end(), kv.second.begin(), kv.second.end());
    }
  }
  ctx->q_groups.clear();
  for (auto &kv : byQ) {
    ctx->q_groups[kv.first] = {(int)T.lists_all.size(), (int)kv.second.size()};
    T.lists_all.insert(T.lists_all.end(), kv.second.begin(), kv.second.end());
  }
  ctx->w_groups.clear();
  for (auto &kv : byW) {
    ctx->w_groups[kv.first] = {(int)T.lists_w.size(), (int)kv.second.size()};
    T.lists_w.insert(T.lists_w.end(), kv.second.begin(), kv.second.end());
  }
  return 0;
}

static int upload_tables(pa_ctx *ctx, const HostTables &T) {
  cudaSetDevice(ctx->device);
  const int nh = (int)ctx->hmms.size();
  auto up = [&](auto &buf, const auto &vec, size_t extra) -> cudaError_t {
    cudaError_t e = buf.ensure(vec.size() + extra);
    if (e != cudaSuccess || vec.empty()) return e;
    return cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice);
  };
  CK(up(ctx->d_packs, ctx->h_packs, 1));
  CK(up(ctx->d_packmem, T.packmem, 1));
  CK(up(ctx->d_hmm_lists_rest, T.lists_rest, 1));
  CK(ctx->d_sort.ensure(3 * ((size_t)nh + 1)));
  CK(up(ctx->d_hmms, ctx->h_dev, 0));
  CK(up(ctx->d_msv, T.msv, 0));
  CK(up(ctx->d_msvh, T.msvh, 1));
  CK(up(ctx->d_v16, T.v16, 4));
  CK(up(ctx->d_vit, T.vit, 0));
  CK(up(ctx->d_fwd, T.fwd, 0));
  CK(up(ctx->d_hmm_lists, T.lists_all, 1));
  CK(up(ctx->d_hmm_lists_w, T.lists_w, 1));
  CK(ctx->d_counters.ensure(16));
  ctx->committed = true;
  return 0;
}

extern "C" int pa_commit_hmms(pa_ctx *ctx) {
  if (!ctx) return -1;
  HostTables T;
  const int rc = build_tables(ctx, T);
  return rc ? rc : upload_tables(ctx, T);
}

// One profile set on several GPUs: the profiles were added to ctxs[0]; the search profiles are configured and the device
// tables laid out ONCE, then uploaded to every context (the others take over ctxs[0]'s profile list and options).
extern "C" int pa_commit_hmms_multi(pa_ctx **ctxs, int n) {
  if (!ctxs || n < 1 || !ctxs[0]) return -1;
  pa_ctx *c0 = ctxs[0];
  HostTables T;
  int rc = build_tables(c0, T);
  if (rc) return rc;
  for (int k = 0; k < n; k++) {
    pa_ctx *c = ctxs[k];
    if (!c) { c0->err = "pa_commit_hmms_multi: null context"; return -1; }
    if (k) {
      c->opts = c0->opts;
      // the other contexts only need what the search reads after the commit (widths, E-value parameters, consensus, scales)
      c->hmms.clear();
      c->hmms.reserve(c0->hmms.size());
      for (const HostProfile &hp : c0->hmms) {
        HostProfile sl;
        sl.name = hp.name; sl.abc = hp.abc; sl.M = hp.M; sl.K = hp.K; sl.Kp = hp.Kp; sl.cons = hp.cons;
        memcpy(sl.ev, hp.ev, sizeof(sl.ev));
        sl.has_stats = hp.has_stats; sl.configured = true;
        sl.tbm_b = hp.tbm_b; sl.tec_b = hp.tec_b; sl.base_b = hp.base_b; sl.bias_b = hp.bias_b;
        sl.scale_b = hp.scale_b; sl.scale_w = hp.scale_w; sl.base_w = hp.base_w; sl.xw_E = hp.xw_E;
        memcpy(sl.eo1, hp.eo1, sizeof(sl.eo1));
        c->hmms.push_back(std::move(sl));
      }
      c->tables_shared = true;
      c->h_dev = c0->h_dev;
      c->v16_classes = c0->v16_classes;
      c->vit_classes = c0->vit_classes;
      c->rank_B = c0->rank_B;
      c->q_groups = c0->q_groups;
      c->q_groups_rest = c0->q_groups_rest;
      c->w_groups = c0->w_groups;
      c->h_packs = c0->h_packs;
      c->maxB = c0->maxB;
      c->maxM = c0->maxM;
      c->maxQ_wide = c0->maxQ_wide;
      c->ftok = c0->ftok;
      c->pack_q = c0->pack_q;
      c->msv_impl = c0->msv_impl;
    }
  }
  // uploads run side by side, one host thread per GPU (pageable-memory copies are staged by the driver per device)
  std::vector<int> rcs(n, 0);
  std::vector<std::thread> th;
  for (int k = 1; k < n; k++) th.emplace_back([&, k] { rcs[k] = upload_tables(ctxs[k], T); });
  rcs[0] = upload_tables(c0, T);
  for (auto &t : th) t.join();
  for (int k = 0; k < n; k++)
    if (rcs[k]) { if (k) c0->err = ctxs[k]->err; return rcs[k]; }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// targets
// ---------------------------------------------------------------------------------------------

// distinct-length tables + lidx after d_tlen is known; h_tlen must be filled
static int build_length_tables(pa_ctx *ctx) {
  const uint32_t n = ctx->ntargets;
  int maxL = 0;
  for (uint32_t i = 0; i < n; i++) maxL = std::max(maxL, (int)ctx->h_tlen[i]);
  ctx->maxL = maxL;
  std::vector<uint8_t> present((size_t)maxL + 1, 0);
  unsigned long long tot = 0;
  for (uint32_t i = 0; i < n; i++) { present[ctx->h_tlen[i]] = 1; tot += ctx->h_tlen[i]; }
  ctx->total_res = tot;
  std::vector<uint16_t> L2li((size_t)maxL + 1, 0);
  ctx->h_L.clear();
  for (int L = 0; L <= maxL; L++)
    if (present[L]) {
      if (ctx->h_L.size() >= 65535) { ctx->err = "more than 65535 distinct target lengths in one chunk: split the chunk"; return -3; }
      L2li[L] = (uint16_t)ctx->h_L.size();
      ctx->h_L.push_back(L);
    }
  const int nL = ctx->nL = (int)ctx->h_L.size();
  std::vector<float> nullsc(nL), bA(nL), bB(nL);
  std::vector<uint8_t> tjb(nL);
  std::vector<int16_t> xw(nL);
  const float scale_b = (float)(3.0 / 0.69314718055994529), scale_w = (float)(500.0 / 0.69314718055994529);
  for (int i = 0; i < nL; i++) {
    const int L = ctx->h_L[i];
    nullsc[i] = null1_score(L);
    tjb[i] = unbiased_byteify(scale_b, logf(3.0f / (float)(L + 3)));
    const float pmove = (2.0f + 1.0f) / ((float)L + 2.0f + 1.0f);
    xw[i] = wordify(scale_w, logf(pmove));
    const float p1 = (float)L / (float)(L + 1);
    bA[i] = (float)L * logf(p1);
    bB[i] = logf((float)(1.0 - (double)p1));
  }
  CK(ctx->d_L.ensure(nL)); CK(ctx->d_nullsc.ensure(nL)); CK(ctx->d_biasA.ensure(nL)); CK(ctx->d_biasB.ensure(nL));
  CK(ctx->d_tjb.ensure(nL)); CK(ctx->d_xwmove.ensure(nL)); CK(ctx->d_L2li.ensure((size_t)maxL + 1));
  CK(cudaMemcpyAsync(ctx->d_L.p, ctx->h_L.data(), nL * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_nullsc.p, nullsc.data(), nL * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_biasA.p, bA.data(), nL * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_biasB.p, bB.data(), nL * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_tjb.p, tjb.data(), nL, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_xwmove.p, xw.data(), nL * 2, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_L2li.p, L2li.data(), ((size_t)maxL + 1) * 2, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx->d_lidx.ensure(n));
  lidx_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_tlen.p, ctx->d_L2li.p, ctx->d_lidx.p, n);
  ctx->stats.kernel_launches++;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int64_t pa_load_reads(pa_ctx *ctx, const char *nt, const uint64_t *off, uint32_t nreads, int moltype,
                                 int gencode, int no_reverse, int codon_only) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  const uint64_t total = off[nreads];
  // capacity limits are -3 ("split the chunk and retry", as for the survivor lists of pa_search); -2 is for bad input
  if (2 * total + 32ull * nreads + 64 >= (1ull << 32)) { ctx->err = "chunk too large (2*nt + 32*reads must stay below 4 GiB): split the chunk"; return -3; }
  int nfwd, nrev;
  if (moltype == PA_TRANSLATE) {
    nfwd = codon_only ? 1 : 3;
    nrev = no_reverse ? 0 : 3;
    ctx->target_abc = PA_AMINO;
  } else {
    nfwd = 1;
    nrev = no_reverse ? 0 : 1;
    ctx->target_abc = PA_DNA;
  }
  const int nframes = nfwd + nrev;
  const uint64_t ntargets64 = (uint64_t)nreads * nframes;
  if (ntargets64 >= (1ull << 32)) { ctx->err = "too many targets in one chunk: split the chunk"; return -3; }
  uint8_t tab64[64];
  memset(tab64, 26, 64);
  if (moltype == PA_TRANSLATE && !codon_table(gencode, tab64)) { ctx->err = "unsupported genetic code table"; return -2; }
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  cudaEventRecord(e0, ctx->stream);
  CK(ctx->d_nt.ensure(total + 16));
  CK(ctx->d_ntoff.ensure((size_t)nreads + 1));
  CK(ctx->d_codon.ensure(64));
  CK(ctx->d_flag.ensure(4));
  CK(cudaMemcpyAsync(ctx->d_nt.p, nt, total, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_ntoff.p, off, ((size_t)nreads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_codon.p, tab64, 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_flag.p, 0, 16, ctx->stream));
  const size_t outbytes = ((2 * total + 3) & ~3ull) + 32ull * nreads + 64;
  CK(ctx->d_dsq.ensure(outbytes));
  CK(ctx->d_toff.ensure(ntargets64 + 1));
  CK(ctx->d_tlen.ensure(ntargets64 + 1));
  cudaEventRecord(e1, ctx->stream);
  const int grid = std::min<uint64_t>((nreads + 7) / 8, (uint64_t)ctx->num_sms * 16);
  if (nreads) {
    if (moltype == PA_TRANSLATE) {
#define PA_TR_LAUNCH(F, R)                                                                                                  \
  translate_kernel<F, R><<<std::max(grid, 1), 256, 0, ctx->stream>>>(ctx->d_nt.p, ctx->d_ntoff.p, nreads, ctx->d_codon.p,   \
                                                                     ctx->d_dsq.p, ctx->d_toff.p, ctx->d_tlen.p, ctx->d_flag.p)
      if (nfwd == 3 && nrev == 3) PA_TR_LAUNCH(3, 3);
      else if (nfwd == 3) PA_TR_LAUNCH(3, 0);
      else if (nrev == 3) PA_TR_LAUNCH(1, 3);
      else PA_TR_LAUNCH(1, 0);
#undef PA_TR_LAUNCH
    }
    else
      translate_generic_kernel<<<std::max(grid, 1), 256, 0, ctx->stream>>>(ctx->d_nt.p, ctx->d_ntoff.p, nreads, 1, nfwd, nrev,
                                                                         ctx->d_codon.p, ctx->d_dsq.p, ctx->d_toff.p, ctx->d_tlen.p,
                                                                         ctx->d_flag.p);
  }
  ctx->stats.kernel_launches++;
  cudaEventRecord(e2, ctx->stream);
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, ctx->d_flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  // host-side target lengths (same arithmetic as the kernel) while the GPU works
  ctx->h_tlen.resize(ntargets64);
  for (uint32_t r = 0; r < nreads; r++) {
    const int L = (int)(off[r + 1] - off[r]);
    for (int f = 0; f < nframes; f++) {
      int len;
      if (moltype == PA_TRANSLATE) {
        const int j = f >= nfwd ? f - nfwd : f;
        len = (L - j) >= 3 ? (L - j) / 3 : 0;
      } else
        len = L;
      ctx->h_tlen[(size_t)r * nframes + f] = (uint32_t)len;
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1); ctx->stats.ms_h2d = ms;
  cudaEventElapsedTime(&ms, e1, e2); ctx->stats.ms_translate = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  if (flag) { ctx->err = "invalid nucleotide character in reads (Biopython would raise TranslationError)"; return -2; }
  ctx->ntargets = (uint32_t)ntargets64;
  ctx->nframes = nframes;
  { const int lrc = build_length_tables(ctx); if (lrc != 0) return lrc; }
  return (int64_t)ntargets64;
}

extern "C" int64_t pa_load_proteins(pa_ctx *ctx, const char *aa, const uint64_t *off, uint32_t nseq, int abc) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  const uint64_t total = off[nseq];
  if (total + 4ull * nseq + 64 >= (1ull << 32)) { ctx->err = "chunk too large: split the chunk"; return -3; }
  std::vector<uint32_t> toff((size_t)nseq + 1);
  ctx->h_tlen.resize(nseq);
  uint32_t pos = 0;
  for (uint32_t i = 0; i < nseq; i++) {
    toff[i] = pos;
    const uint32_t L = (uint32_t)(off[i + 1] - off[i]);
    ctx->h_tlen[i] = L;
    pos += (L + 3) & ~3u;
  }
  toff[nseq] = pos;
  CK(ctx->d_nt.ensure(total + 16));
  CK(ctx->d_ntoff.ensure((size_t)nseq + 1));
  CK(ctx->d_dsq.ensure((size_t)pos + 64));
  CK(ctx->d_toff.ensure((size_t)nseq + 1));
  CK(ctx->d_tlen.ensure((size_t)nseq + 1));
  CK(cudaMemcpyAsync(ctx->d_nt.p, aa, total, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_ntoff.p, off, ((size_t)nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_toff.p, toff.data(), ((size_t)nseq + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_tlen.p, ctx->h_tlen.data(), (size_t)nseq * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_dsq.p, 0, (size_t)pos + 64, ctx->stream));
  if (nseq) {
    const int grid = std::min<uint64_t>((nseq + 7) / 8, (uint64_t)ctx->num_sms * 8);
    digitize_kernel<<<std::max(grid, 1), 256, 0, ctx->stream>>>(ctx->d_nt.p, ctx->d_ntoff.p, ctx->d_toff.p, nseq, abc, ctx->d_dsq.p);
    ctx->stats.kernel_launches++;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  ctx->ntargets = nseq;
  ctx->nframes = 1;
  ctx->target_abc = abc;
  { const int lrc = build_length_tables(ctx); if (lrc != 0) return lrc; }
  return nseq;
}

extern "C" int pa_num_frames(pa_ctx *ctx) { return ctx ? ctx->nframes : -1; }

extern "C" int64_t pa_get_targets(pa_ctx *ctx, char *out, uint64_t cap, uint64_t *out_off) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  const uint32_t n = ctx->ntargets;
  std::vector<uint32_t> toff(n);
  if (n) CK(cudaMemcpy(toff.data(), ctx->d_toff.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  size_t bytes = 0;
  for (uint32_t i = 0; i < n; i++) bytes = std::max<size_t>(bytes, (size_t)toff[i] + ctx->h_tlen[i]);
  std::vector<uint8_t> dsq(bytes + 1);
  if (bytes) CK(cudaMemcpy(dsq.data(), ctx->d_dsq.p, bytes, cudaMemcpyDeviceToHost));
  const char *syms = Alphabet::get(ctx->target_abc).syms;
  uint64_t pos = 0;
  for (uint32_t i = 0; i < n; i++) {
    out_off[i] = pos;
    if (pos + ctx->h_tlen[i] > cap) { ctx->err = "pa_get_targets: output buffer too small"; return -3; }
    for (uint32_t j = 0; j < ctx->h_tlen[i]; j++) out[pos + j] = syms[dsq[toff[i] + j]];
    pos += ctx->h_tlen[i];
  }
  out_off[n] = pos;
  return (int64_t)n;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
static cudaError_t dispatch_msv(pa_ctx *ctx, int Q, MsvArgs &a, int nrows) {
  const bool tmem = msv_use_tmem(ctx, Q, nrows);
  if (tmem) a.tab = ctx->d_msvh.p;
  return launch_msv_q(Q, tmem, tmem && msv_fullt(ctx, Q), a, nrows, ctx->num_sms, ctx->stream, &ctx->stats.kernel_launches);
}

static cudaError_t side_fork(pa_ctx *ctx);
static cudaError_t side_join(pa_ctx *ctx);

static cudaError_t dispatch_vit16(pa_ctx *ctx, int BH, int NW, Vit16Args &v, cudaStream_t st) {
  return launch_vit16_class(BH, NW, v, ctx->num_sms, st, &ctx->stats.kernel_launches);
}
static cudaError_t dispatch_vit(pa_ctx *ctx, int Bv, int NW, VitArgs &v, cudaStream_t st) {
  return launch_vit_class(Bv, NW, v, ctx->num_sms, st, &ctx->stats.kernel_launches);
}

// Fork: side streams wait for everything queued on the main stream; join: the main stream waits for them.
// Used to run the per-class launches of one stage concurrently (few, long pairs per class otherwise serialise).
static cudaError_t side_fork(pa_ctx *ctx) {
  for (int i = 0; i < 8; i++)
    if (!ctx->side[i]) { cudaError_t e = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
  for (int i = 0; i < 9; i++)
    if (!ctx->side_ev[i]) { cudaError_t e = cudaEventCreateWithFlags(&ctx->side_ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  cudaError_t e = cudaEventRecord(ctx->side_ev[8], ctx->stream);
  if (e != cudaSuccess) return e;
  for (int i = 0; i < 8; i++) { e = cudaStreamWaitEvent(ctx->side[i], ctx->side_ev[8], 0); if (e != cudaSuccess) return e; }
  return cudaSuccess;
}
static cudaError_t side_join(pa_ctx *ctx) {
  for (int i = 0; i < 8; i++) {
    cudaError_t e = cudaEventRecord(ctx->side_ev[i], ctx->side[i]);
    if (e != cudaSuccess) return e;
    e = cudaStreamWaitEvent(ctx->stream, ctx->side_ev[i], 0);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

static cudaError_t dispatch_fwd_t(pa_ctx *ctx, int B, FwdTArgs &a, cudaStream_t st) {
  float *scratch = nullptr;
  if (B > 64 && a.end > a.begin) {  // fwd_wide_kernel: one row of P per CTA, sized for the widest profile; the launches of the
    // size classes run side by side on different streams, so each class has its own region
    const size_t region = fwd_wide_per_cta(ctx->maxB) * (size_t)fwd_wide_grid_cap(ctx->num_sms);
    cudaError_t e = ctx->d_fwdw.ensure(region * 4);
    if (e != cudaSuccess) return e;
    const int cls = fwd_wide_class(B);
    scratch = ctx->d_fwdw.p + region * (size_t)(cls == 96 ? 0 : cls == 128 ? 1 : cls == 192 ? 2 : 3);
  }
  return launch_fwd_b(B, a, ctx->num_sms, scratch, ctx->maxB, st, &ctx->stats.kernel_launches);
}

// counting sort of n Pass2 records by profile rank: in -> out; start[] (host copy) gets nh + 1 offsets
static int sort_by_profile(pa_ctx *ctx, const Pass2 *in, Pass2 *out, unsigned long long n, std::vector<unsigned int> &start) {
  const int nh = (int)ctx->hmms.size();
  unsigned int *cnt = ctx->d_sort.p, *st = cnt + (nh + 1), *cur = st + (nh + 1);
  CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * (nh + 1), ctx->stream));
  hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, n, ctx->d_hmms.p, cnt);
  scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, nh, st, cur);
  scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, n, ctx->d_hmms.p, cur, out);
  ctx->stats.kernel_launches += 3;
  start.resize((size_t)nh + 1);
  CK(cudaMemcpyAsync(start.data(), st, sizeof(unsigned int) * (nh + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int launch_thresholds(pa_ctx *ctx, int all_pass) {
  const int nh = (int)ctx->hmms.size();
  CK(ctx->d_thrJ.ensure((size_t)nh * ctx->nL));
  CK(ctx->d_scJ.ensure((size_t)nh * ctx->nL));
  const int n = nh * ctx->nL;
  msv_threshold_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_hmms.p, nh, make_lentabs(ctx), ctx->d_thrJ.p, ctx->d_scJ.p, all_pass);
  ctx->stats.kernel_launches++;
  return 0;
}

static int check_targets_ready(pa_ctx *ctx) {
  if (!ctx->committed) { ctx->err = "profiles not committed (call pa_commit_hmms)"; return -2; }
  if (ctx->ntargets == 0) return 0;
  if (ctx->target_abc != ctx->hmms[0].abc) { ctx->err = "target alphabet does not match the profiles"; return -2; }
  return 0;
}

#include "search_impl.inc"

// host-side view of the MSV work-item order (same function the kernels run); used by the CPU test-suite
extern "C" int pa_msv_item_decode(uint64_t item, uint32_t nlist, uint32_t ntiles, uint32_t super_tiles, uint32_t *slot, uint32_t *tile) {
  if (!slot || !tile || nlist == 0 || ntiles == 0 || super_tiles == 0 || item >= (uint64_t)nlist * ntiles) return -2;
  msv_item_decode(item, nlist, ntiles, super_tiles, *slot, *tile);
  return 0;
}

// accessors for the other translation units (ctx_access.hpp)
cudaStream_t pa_ctx_stream(pa_ctx *ctx) { return ctx->stream; }
int pa_ctx_device(pa_ctx *ctx) { return ctx->device; }
int pa_ctx_num_sms(pa_ctx *ctx) { return ctx->num_sms; }
void pa_ctx_set_error(pa_ctx *ctx, const std::string &msg) { ctx->err = msg; }
void pa_ctx_count_launches(pa_ctx *ctx, int n) { ctx->stats.kernel_launches += n; }

