// mapping.cuh -- hit -> reference-column mapping on the GPU (HBM-bound byte work).
//
// One thread per hit. Restates, for the default settings (extend_pos = 0, no -b pre-filter):
//   pretreat_hits()      /root/reference/lib/assemble.py:767-873  (terminal trimming by column frequency)
//   read_hit.__init__()  /root/reference/lib/assemble.py:13-81, 151-202 (column walk: codon / base per
//                        reference column, unmapped insert positions, read interval)
// Inputs are the PhyloAln-form strings of hit_info.tsv (upper case, '-' gaps), the raw read, and the
// per-column residue counts of the reference alignment (get_ref_score's site_composition).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/phyloaln_b200.h"

namespace pa {

#define PA_MAP_UNROLL 1

__device__ __forceinline__ int comp_sym(char c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a';
  if (c == '-') return 26;
  if (c == '*') return 27;
  return 28;
}
__device__ __forceinline__ char nt_upper(char c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
__device__ __forceinline__ char nt_comp(char c) {
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'M': return 'K'; case 'R': return 'Y'; case 'W': return 'W'; case 'S': return 'S';
    case 'Y': return 'R'; case 'K': return 'M'; case 'V': return 'B'; case 'H': return 'D';
    case 'D': return 'H'; case 'B': return 'V'; case 'X': return 'X'; case 'N': return 'N';
    default: return c;
  }
}

struct MapArgs {
  int nhits;
  const char *model, *aligned;
  const long long *ali_off;
  const char *reads;
  const long long *read_off;
  const int *hmm_from, *hmm_to, *ali_from, *ali_to, *frame, *strand;
  const int *comp;
  const long long *comp_off;
  const int *ncols_ref;
  int trim_pos;
  double pos_freq;
  pa_maprec *rec;
  char *base_pool, *codon_pool;
  int *unmap_pool;
  const long long *col_off, *unmap_off;   // per-hit reservations (host prefix sums)
};

__device__ __forceinline__ bool low_freq(const MapArgs &a, int h, int pos, char c) {
  if (pos < 1 || pos > a.ncols_ref[h]) return true;
  const int *row = a.comp + ((size_t)a.comp_off[h] + pos - 1) * 32;
  int tot = 0;
  for (int s = 0; s < 32; s++) tot += row[s];
  const int sy = comp_sym(c);
  const int cnt = sy < 32 ? row[sy] : 0;
  if (tot == 0) return true;
  return (double)cnt / (double)tot < a.pos_freq;
}

__global__ void map_kernel(MapArgs a) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= a.nhits) return;
  const char *mo = a.model + a.ali_off[h], *al = a.aligned + a.ali_off[h];
  int n = (int)(a.ali_off[h + 1] - a.ali_off[h]);
  const char *rd = a.reads + a.read_off[h];
  const int rlen = (int)(a.read_off[h + 1] - a.read_off[h]);
  pa_maprec r;
  r.keep = 0; r.qstart = r.qend = 0; r.ali_from = a.ali_from[h]; r.ali_to = a.ali_to[h];
  r.start_trim = r.end_trim = 0; r.read_from = r.read_to = 0; r.strand = a.strand[h];
  r.col_off = (int)a.col_off[h]; r.ncols = 0; r.unmap_off = (int)a.unmap_off[h]; r.n_unmap = 0;
  int hmm_from = a.hmm_from[h], hmm_to = a.hmm_to[h];
  // ---- pretreat_hits: short hits, terminal trimming
  if (n <= 3) { a.rec[h] = r; return; }
  int st = 0, i = 0, j = 0;
  while (st < a.trim_pos && i < n) {
    if (mo[i] != '-') {
      if (low_freq(a, h, hmm_from + st, al[i])) st++;
      else break;
    }
    if (al[i] != '-') j++;
    i++;
  }
  if (i == n) { a.rec[h] = r; return; }
  while (i < n && al[i] == '-') { st++; i++; }
  hmm_from += st;
  mo += i; al += i; n -= i;
  r.ali_from += j;
  int et = 0;
  i = n - 1; j = 0;
  while (et < a.trim_pos && i >= 0) {
    if (mo[i] != '-') {
      if (low_freq(a, h, hmm_to - et, al[i])) et++;
      else break;
    }
    if (al[i] != '-') j++;
    i--;
  }
  if (i < 0) { a.rec[h] = r; return; }
  while (i >= 0 && al[i] == '-') { et++; i--; }
  hmm_to -= et;
  n = i + 1;
  r.ali_to -= j;
  r.start_trim = st; r.end_trim = et;
  r.qstart = hmm_from; r.qend = hmm_to;
  // ---- read_hit.__init__: nucleotide slice + column walk
  const int fr = a.frame[h];
  const bool rev = a.strand[h] < 0;
  const int step = fr >= 0 ? 3 : 1;
  int s0, s1;  // 0-based half-open slice on the (possibly reverse-complemented) read
  if (fr >= 0) { s0 = 3 * r.ali_from + fr - 3; s1 = 3 * r.ali_to + fr; }
  else { s0 = r.ali_from - 1; s1 = r.ali_to; }
  if (s0 < 0) s0 = 0;
  if (s1 > rlen) s1 = rlen;
  const int slen = s1 > s0 ? s1 - s0 : 0;
  int SE0, SE1;
  if (fr >= 0) {
    if (!rev) { SE0 = 3 * r.ali_from + fr - 2; SE1 = 3 * r.ali_to + fr; }
    else { SE0 = rlen - 3 * r.ali_to - fr + 1; SE1 = rlen - 3 * r.ali_from - fr + 3; }
  } else {
    if (!rev) { SE0 = r.ali_from; SE1 = r.ali_to; }
    else { SE0 = rlen - r.ali_to + 1; SE1 = rlen - r.ali_from + 1; }
  }
  r.read_from = SE0; r.read_to = SE1;
  char *bp = a.base_pool + a.col_off[h];
  char *cp = a.codon_pool + 3 * a.col_off[h];
  int *up = a.unmap_pool + a.unmap_off[h];
  int pos = 0, col = 0, nun = 0;
  for (int q = 0; q + step <= slen || (step == 1 && q < slen); q += step) {
    while (pos < n && al[pos] == '-') {
      bp[col] = '-';
      if (fr >= 0) { cp[3 * col] = '-'; cp[3 * col + 1] = '-'; cp[3 * col + 2] = '-'; }
      col++; pos++;
    }
    if (pos >= n) break;
    char nt3[3];
    for (int k = 0; k < step; k++) {
      const int idx = s0 + q + k;  // index on the oriented read
      char c = rev ? nt_comp(nt_upper(rd[rlen - 1 - idx])) : nt_upper(rd[idx]);
      nt3[k] = c;
    }
    if (mo[pos] == '-') {
      for (int k = 0; k < step; k++) up[nun++] = rev ? SE1 - q - k : SE0 + q + k;
    } else {
      if (fr >= 0) { cp[3 * col] = nt3[0]; cp[3 * col + 1] = nt3[1]; cp[3 * col + 2] = nt3[2]; bp[col] = al[pos]; }
      else bp[col] = nt3[0];
      col++;
    }
    pos++;
  }
  r.ncols = col;
  r.n_unmap = nun;
  r.keep = 1;
  a.rec[h] = r;
}

}  // namespace pa
