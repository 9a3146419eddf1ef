// mapping.cu -- host side of pa_map_hits / pa_map_hits_ex: reservations (prefix sums), H2D, map_kernel
// (+ accum_kernel), D2H.  Kernels in mapping.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/phyloaln_b200.h"
#include "ctx_access.hpp"
#include "gencode.hpp"
#include "mapping.cuh"

using namespace pa;

namespace {
struct DevPool {  // frees everything on scope exit
  std::vector<void *> ptrs;
  ~DevPool() { for (void *p : ptrs) cudaFree(p); }
  template <class T>
  cudaError_t alloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};
}  // namespace

extern "C" int pa_map_hits_ex(pa_ctx *ctx, int32_t nhits, const char *model, const char *aligned, const int64_t *ali_off,
                              const char *reads, const int64_t *read_off, const int32_t *hmm_from, const int32_t *hmm_to,
                              const int32_t *ali_from, const int32_t *ali_to, const int32_t *frame, const int32_t *strand,
                              const int32_t *comp, const int64_t *comp_off, const int32_t *ncols_ref, const double *out_score,
                              const int64_t *out_off, const int32_t *n_out, const pa_map_opts *opts, const pa_accum *acc,
                              pa_maprec *rec, char *base_pool, char *codon_pool, int32_t *unmap_pool, int64_t pool_cols_cap,
                              int64_t unmap_cap) {
  if (!ctx) return -1;
  if (!opts) { pa_ctx_set_error(ctx, "pa_map_hits_ex: opts is NULL"); return -2; }
  cudaSetDevice(pa_ctx_device(ctx));
  if (nhits <= 0) return 0;
  const int ext = std::max(0, opts->extend_pos);
  const bool prefilter = opts->prefilter && out_score && out_off && n_out;
  uint8_t tab64[64];
  if (!codon_table(opts->gencode, tab64)) { pa_ctx_set_error(ctx, "pa_map_hits_ex: unsupported genetic code table"); return -2; }
  std::vector<long long> col_off((size_t)nhits + 1), un_off((size_t)nhits + 1);
  long long c = 0, u = 0, comp_rows = 0, out_elems = 0;
  for (int h = 0; h < nhits; h++) {
    col_off[h] = c;
    un_off[h] = u;
    c += std::max(0, hmm_to[h] - hmm_from[h] + 1) + 1 + 2 * ext;
    u += 3 * (ali_off[h + 1] - ali_off[h]) + 3;
    comp_rows = std::max<long long>(comp_rows, comp_off[h] + ncols_ref[h]);
    if (prefilter) out_elems = std::max<long long>(out_elems, out_off[h] + (long long)n_out[h] * ncols_ref[h]);
  }
  col_off[nhits] = c;
  un_off[nhits] = u;
  if (c > pool_cols_cap || u > unmap_cap) { pa_ctx_set_error(ctx, "pa_map_hits: output pools too small"); return -3; }
  if (c >= INT_MAX || u >= INT_MAX) { pa_ctx_set_error(ctx, "pa_map_hits: too many columns in one call"); return -3; }
  const size_t nali = (size_t)ali_off[nhits], nrd = (size_t)read_off[nhits], N = (size_t)nhits;
  DevPool pool;
  char *d_model, *d_al, *d_reads, *d_base, *d_codon;
  long long *d_alioff, *d_rdoff, *d_compoff, *d_coloff, *d_unoff, *d_outoff = nullptr, *d_binrow = nullptr;
  int *d_i32, *d_comp, *d_unmap, *d_nout = nullptr, *d_bin = nullptr, *d_rep = nullptr, *d_hist = nullptr, *d_first = nullptr,
      *d_qrange = nullptr;
  double *d_out = nullptr;
  uint8_t *d_tab;
  pa_maprec *d_rec;
#define MCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { pa_ctx_set_error(ctx, std::string("pa_map_hits: ") + cudaGetErrorString(e_)); return -1; } } while (0)
  MCK(pool.alloc(&d_model, nali)); MCK(pool.alloc(&d_al, nali)); MCK(pool.alloc(&d_reads, nrd));
  MCK(pool.alloc(&d_base, (size_t)c)); MCK(pool.alloc(&d_codon, 3 * (size_t)c));
  MCK(pool.alloc(&d_alioff, N + 1)); MCK(pool.alloc(&d_rdoff, N + 1)); MCK(pool.alloc(&d_compoff, N));
  MCK(pool.alloc(&d_coloff, N + 1)); MCK(pool.alloc(&d_unoff, N + 1)); MCK(pool.alloc(&d_i32, 7 * N));
  MCK(pool.alloc(&d_comp, 32 * (size_t)comp_rows)); MCK(pool.alloc(&d_unmap, (size_t)u)); MCK(pool.alloc(&d_rec, N));
  MCK(pool.alloc(&d_tab, 64));
  cudaStream_t s = pa_ctx_stream(ctx);
  MCK(cudaMemcpyAsync(d_model, model, nali, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_al, aligned, nali, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_reads, reads, nrd, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_alioff, ali_off, 8 * (N + 1), cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_rdoff, read_off, 8 * (N + 1), cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_compoff, comp_off, 8 * N, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_coloff, col_off.data(), 8 * (N + 1), cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_unoff, un_off.data(), 8 * (N + 1), cudaMemcpyHostToDevice, s));
  const int32_t *i32src[7] = {hmm_from, hmm_to, ali_from, ali_to, frame, strand, ncols_ref};
  for (int k = 0; k < 7; k++) MCK(cudaMemcpyAsync(d_i32 + (size_t)k * N, i32src[k], 4 * N, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_comp, comp, 4 * 32 * (size_t)comp_rows, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_tab, tab64, 64, cudaMemcpyHostToDevice, s));
  if (prefilter) {
    MCK(pool.alloc(&d_out, (size_t)out_elems)); MCK(pool.alloc(&d_outoff, N)); MCK(pool.alloc(&d_nout, N));
    MCK(cudaMemcpyAsync(d_out, out_score, 8 * (size_t)out_elems, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(d_outoff, out_off, 8 * N, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(d_nout, n_out, 4 * N, cudaMemcpyHostToDevice, s));
  }
  MapArgs a;
  a.nhits = nhits; a.model = d_model; a.aligned = d_al; a.ali_off = d_alioff; a.reads = d_reads; a.read_off = d_rdoff;
  a.hmm_from = d_i32; a.hmm_to = d_i32 + N; a.ali_from = d_i32 + 2 * N; a.ali_to = d_i32 + 3 * N;
  a.frame = d_i32 + 4 * N; a.strand = d_i32 + 5 * N; a.ncols_ref = d_i32 + 6 * N;
  a.comp = d_comp; a.comp_off = d_compoff; a.trim_pos = opts->trim_pos; a.extend_pos = ext; a.prefilter = prefilter ? 1 : 0;
  a.pos_freq = opts->pos_freq; a.outgroup_weight = opts->outgroup_weight; a.tab64 = d_tab;
  a.out_score = d_out; a.out_off = d_outoff; a.n_out = d_nout; a.rec = d_rec;
  a.base_pool = d_base; a.codon_pool = d_codon; a.unmap_pool = d_unmap; a.col_off = d_coloff; a.unmap_off = d_unoff;
  map_kernel<<<(nhits + 127) / 128, 128, 0, s>>>(a);
  pa_ctx_count_launches(ctx, 1);
  MCK(cudaGetLastError());
  size_t cells = 0;
  if (acc) {
    if (!acc->bin || !acc->rep || !acc->bin_row || !acc->hist || !acc->first || !acc->qrange || acc->nbins <= 0) {
      pa_ctx_set_error(ctx, "pa_map_hits_ex: incomplete accumulation request");
      return -2;
    }
    cells = (size_t)acc->bin_row[acc->nbins] * 96;
    MCK(pool.alloc(&d_bin, N)); MCK(pool.alloc(&d_rep, N)); MCK(pool.alloc(&d_binrow, (size_t)acc->nbins + 1));
    MCK(pool.alloc(&d_hist, cells)); MCK(pool.alloc(&d_first, cells)); MCK(pool.alloc(&d_qrange, 2 * (size_t)acc->nbins));
    MCK(cudaMemcpyAsync(d_bin, acc->bin, 4 * N, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(d_rep, acc->rep, 4 * N, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(d_binrow, acc->bin_row, 8 * ((size_t)acc->nbins + 1), cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(d_qrange, acc->qrange, 8 * (size_t)acc->nbins, cudaMemcpyHostToDevice, s));
    MCK(cudaMemsetAsync(d_hist, 0, 4 * cells, s));
    MCK(cudaMemsetAsync(d_first, 0x7f, 4 * cells, s));  // 0x7f7f7f7f: larger than any hit index
    AccumArgs q;
    q.nhits = nhits; q.rec = d_rec; q.base_pool = d_base; q.codon_pool = d_codon; q.frame = a.frame; q.bin = d_bin; q.rep = d_rep;
    q.bin_row = d_binrow; q.hist = d_hist; q.first = d_first; q.qrange = d_qrange;
    accum_kernel<<<(unsigned)((N * 32 + 255) / 256), 256, 0, s>>>(q);
    pa_ctx_count_launches(ctx, 1);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(acc->hist, d_hist, 4 * cells, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(acc->first, d_first, 4 * cells, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(acc->qrange, d_qrange, 8 * (size_t)acc->nbins, cudaMemcpyDeviceToHost, s));
  }
  MCK(cudaMemcpyAsync(rec, d_rec, sizeof(pa_maprec) * N, cudaMemcpyDeviceToHost, s));
  MCK(cudaMemcpyAsync(base_pool, d_base, (size_t)c, cudaMemcpyDeviceToHost, s));
  MCK(cudaMemcpyAsync(codon_pool, d_codon, 3 * (size_t)c, cudaMemcpyDeviceToHost, s));
  MCK(cudaMemcpyAsync(unmap_pool, d_unmap, 4 * (size_t)u, cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
  if (acc)
    for (size_t i = 0; i < cells; i++)
      if (acc->first[i] == 0x7f7f7f7f) acc->first[i] = INT_MAX;
#undef MCK
  return 0;
}

extern "C" int pa_map_hits(pa_ctx *ctx, int32_t nhits, const char *model, const char *aligned, const int64_t *ali_off,
                           const char *reads, const int64_t *read_off, const int32_t *hmm_from, const int32_t *hmm_to,
                           const int32_t *ali_from, const int32_t *ali_to, const int32_t *frame, const int32_t *strand,
                           const int32_t *comp, const int64_t *comp_off, const int32_t *ncols_ref, int trim_pos,
                           double pos_freq, pa_maprec *rec, char *base_pool, char *codon_pool, int32_t *unmap_pool,
                           int64_t pool_cols_cap, int64_t unmap_cap) {
  pa_map_opts o;
  o.trim_pos = trim_pos; o.extend_pos = 0; o.gencode = 1; o.prefilter = 0; o.pos_freq = pos_freq; o.outgroup_weight = 1.0;
  return pa_map_hits_ex(ctx, nhits, model, aligned, ali_off, reads, read_off, hmm_from, hmm_to, ali_from, ali_to, frame, strand, comp,
                        comp_off, ncols_ref, nullptr, nullptr, nullptr, &o, nullptr, rec, base_pool, codon_pool, unmap_pool,
                        pool_cols_cap, unmap_cap);
}

// merge_hits() histograms as a segmented fold on the GPU (see merge_kernel, mapping.cuh)
extern "C" int pa_merge_hists(pa_ctx *ctx, int32_t nchains, const int32_t *chain_off, const int32_t *item_q0, const int32_t *item_ncols,
                              const int64_t *item_row, const int32_t *hist_in, const int32_t *key_in, const int64_t *chain_row,
                              const int32_t *chain_ncols, int32_t *hist_out, int32_t *who_out, int32_t *key_out) {
  if (!ctx) return -1;
  if (nchains <= 0) return 0;
  cudaSetDevice(pa_ctx_device(ctx));
  cudaStream_t s = pa_ctx_stream(ctx);
  const int nitems = chain_off[nchains];
  long long in_rows = 0, out_rows = 0;
  int maxcols = 1;
  for (int i = 0; i < nitems; i++) {
    if (item_ncols[i] < 0 || item_row[i] < 0) { pa_ctx_set_error(ctx, "pa_merge_hists: bad item table"); return -2; }
    in_rows = std::max<long long>(in_rows, item_row[i] + item_ncols[i]);
  }
  for (int j = 0; j < nchains; j++) {
    if (chain_off[j + 1] < chain_off[j] || chain_ncols[j] < 0) { pa_ctx_set_error(ctx, "pa_merge_hists: bad chain table"); return -2; }
    out_rows = std::max<long long>(out_rows, chain_row[j] + chain_ncols[j]);
    maxcols = std::max(maxcols, (int)chain_ncols[j]);
  }
  DevPool pool;
#define MCK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) { pa_ctx_set_error(ctx, std::string("pa_merge_hists: ") + cudaGetErrorString(e_)); return -1; } \
  } while (0)
  int *d_coff, *d_q0, *d_nc, *d_cnc, *d_hin, *d_kin, *d_hout, *d_who, *d_kout;
  long long *d_irow, *d_crow;
  const size_t NI = (size_t)nitems, NC = (size_t)nchains, CI = (size_t)in_rows * 96, CO = (size_t)out_rows * 96;
  MCK(pool.alloc(&d_coff, NC + 1)); MCK(pool.alloc(&d_q0, NI)); MCK(pool.alloc(&d_nc, NI)); MCK(pool.alloc(&d_irow, NI));
  MCK(pool.alloc(&d_crow, NC)); MCK(pool.alloc(&d_cnc, NC)); MCK(pool.alloc(&d_hin, CI)); MCK(pool.alloc(&d_kin, CI));
  MCK(pool.alloc(&d_hout, CO)); MCK(pool.alloc(&d_who, CO)); MCK(pool.alloc(&d_kout, CO));
  MCK(cudaMemcpyAsync(d_coff, chain_off, 4 * (NC + 1), cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_q0, item_q0, 4 * NI, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_nc, item_ncols, 4 * NI, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_irow, item_row, 8 * NI, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_crow, chain_row, 8 * NC, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_cnc, chain_ncols, 4 * NC, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_hin, hist_in, 4 * CI, cudaMemcpyHostToDevice, s));
  MCK(cudaMemcpyAsync(d_kin, key_in, 4 * CI, cudaMemcpyHostToDevice, s));
  MergeArgs m;
  m.nchains = nchains; m.chain_off = d_coff; m.item_q0 = d_q0; m.item_ncols = d_nc; m.item_row = d_irow; m.chain_row = d_crow;
  m.chain_ncols = d_cnc; m.hist_in = d_hin; m.key_in = d_kin; m.hist_out = d_hout; m.who_out = d_who; m.key_out = d_kout;
  const unsigned gx = (unsigned)std::min<long long>(((long long)maxcols * 96 + 255) / 256, 1024);
  for (int j0 = 0; j0 < nchains; j0 += 65535) {   // gridDim.y limit
    MergeArgs mm = m;
    const int nj = std::min(65535, nchains - j0);
    mm.chain_off += j0; mm.chain_row += j0; mm.chain_ncols += j0;
    // chain_off values stay absolute item indices; `who` is relative to the chain's first item
    merge_kernel<<<dim3(gx, (unsigned)nj), 256, 0, s>>>(mm);
    pa_ctx_count_launches(ctx, 1);
    MCK(cudaGetLastError());
  }
  MCK(cudaMemcpyAsync(hist_out, d_hout, 4 * CO, cudaMemcpyDeviceToHost, s));
  MCK(cudaMemcpyAsync(who_out, d_who, 4 * CO, cudaMemcpyDeviceToHost, s));
  MCK(cudaMemcpyAsync(key_out, d_kout, 4 * CO, cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
#undef MCK
  return 0;
}
