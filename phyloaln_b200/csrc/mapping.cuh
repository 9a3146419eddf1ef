// mapping.cuh -- hit -> reference-column mapping on the GPU (HBM-bound byte work).
//
// map_kernel, one thread per hit, restates
//   pretreat_hits()      /root/reference/lib/assemble.py:767-873  (short-hit drop, the -b outgroup-score
//                        pre-filter :771-798, terminal trimming by column frequency :800-869)
//   read_hit.__init__()  /root/reference/lib/assemble.py:13-306  (column walk: codon / base per reference
//                        column, unmapped insert positions, read interval; --extend_pos terminal
//                        extension :82-150 nucleotide, :203-298 codon) for the strands the hmmsearch
//                        step produces ('+' and 'rev').
// accum_kernel, one warp per kept hit, is the accumulation of read_hit.consensus()
//   (/root/reference/lib/assemble.py:309-633): per (alignment, species) bin, per reference column (and
//   codon position), counts of each nucleotide weighted by the number of merged duplicate reads, the
//   bin's column range, and the index of the first hit that contributed each symbol (Python dict
//   insertion order, which breaks ties in most_common()).
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

// IUPAC nucleotide -> bit mask T1 C2 A4 G8 (0 = not a nucleotide letter), upper case input
__device__ __forceinline__ int map_nt_mask(char c) {
  switch (c) {
    case 'T': case 'U': return 1; case 'C': return 2; case 'A': return 4; case 'G': return 8;
    case 'R': return 12; case 'Y': return 3; case 'M': return 6; case 'K': return 9; case 'S': return 10;
    case 'W': return 5; case 'H': return 7; case 'B': return 11; case 'V': return 14; case 'D': return 13;
    case 'N': return 15; default: return 0;
  }
}
// Bio.Seq.translate() of one codon: 0 if a letter is not a nucleotide (Biopython raises)
__device__ __noinline__ char map_translate(char c0, char c1, char c2, const uint8_t *tab64) {
  const char *syms = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*";
  const int m0 = map_nt_mask(c0), m1 = map_nt_mask(c1), m2 = map_nt_mask(c2);
  if (!m0 || !m1 || !m2) return 0;
  uint32_t seen = 0;
  int nstop = 0, ntot = 0;
  for (int x = m0; x; x &= x - 1)
    for (int y = m1; y; y &= y - 1)
      for (int z = m2; z; z &= z - 1) {
        const int r = tab64[(__ffs(x) - 1) * 16 + (__ffs(y) - 1) * 4 + (__ffs(z) - 1)];
        ntot++;
        if (r == 27) nstop++;
        else seen |= 1u << r;
      }
  int code;
  if (nstop == ntot) code = 27;
  else if (nstop > 0) code = 26;
  else if (__popc(seen) == 1) code = __ffs(seen) - 1;
  else if ((seen & ~((1u << 2) | (1u << 11))) == 0) code = 21;  // D,N -> B
  else if ((seen & ~((1u << 3) | (1u << 13))) == 0) code = 23;  // E,Q -> Z
  else if ((seen & ~((1u << 7) | (1u << 9))) == 0) code = 22;   // I,L -> J
  else code = 26;
  return syms[code];
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
  int trim_pos, extend_pos, prefilter;
  double pos_freq, outgroup_weight;
  const uint8_t *tab64;                   // genetic code of --gencode (extension translates flanking codons)
  const double *out_score;                // outgroup_score rows: NaN = no entry for that column
  const long long *out_off;               // per hit: first element of its alignment's table
  const int *n_out;                       // per hit: rows (outgroup sequences + PhyloAln_Alt), each ncols_ref[h] long
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
  r.ext_start = r.ext_end = 0;
  int hmm_from = a.hmm_from[h], hmm_to = a.hmm_to[h];
  // ---- pretreat_hits: short hits, terminal trimming
  if (n <= 3) { a.rec[h] = r; return; }
  // ---- pretreat_hits :771-798: conservative score of the hit vs the outgroups over the same columns
  if (a.prefilter && a.n_out[h] > 0) {
    const int nc = a.ncols_ref[h];
    long long score = 0;
    int jm = 0;
    for (int p = 0; p < n; p++)
      if (mo[p] != '-') {
        const int pos = hmm_from + jm;
        if (pos >= 1 && pos <= nc) {
          const int sy = comp_sym(al[p]);
          score += a.comp[((size_t)a.comp_off[h] + pos - 1) * 32 + sy];
        }
        jm++;
      }
    double best = 0.0;
    bool any = false;
    for (int s = 0; s < a.n_out[h]; s++) {
      const double *row = a.out_score + a.out_off[h] + (size_t)s * nc;
      if (hmm_from < 1 || hmm_to > nc || hmm_from > hmm_to) continue;
      bool ok = true;
      for (int pos = hmm_from; pos <= hmm_to && ok; pos++) ok = !isnan(row[pos - 1]);
      if (!ok) continue;
      double sum = 0.0;
      for (int pos = hmm_from; pos < hmm_from + jm && pos <= nc; pos++) sum = __dadd_rn(sum, row[pos - 1]);
      if (!any || sum < best) best = sum;
      any = true;
    }
    if (any && (double)score < __dmul_rn(best, a.outgroup_weight)) { a.rec[h] = r; return; }
  }
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
  char *bp = a.base_pool + a.col_off[h] + a.extend_pos;        // extend_pos columns reserved in front
  char *cp = a.codon_pool + 3 * (a.col_off[h] + a.extend_pos);
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
  // ---- read_hit.__init__ :82-150 / :203-298: extend untrimmed ends while the flanking residue is frequent
  //      enough in the reference column
  auto oriented = [&](int idx) -> char { return rev ? nt_comp(nt_upper(rd[rlen - 1 - idx])) : nt_upper(rd[idx]); };
  int es = 0, ee = 0;
  if (st == 0) {
    for (int e = 1; e <= a.extend_pos && r.qstart - e > 0; e++) {
      if (r.ali_from - e < 1) break;
      char b, c3[3] = {0, 0, 0};
      if (fr >= 0) {
        const int lo = 3 * r.ali_from + fr - 3 * e - 3;
        if (lo < 0 || lo + 3 > rlen) break;
        for (int k = 0; k < 3; k++) c3[k] = oriented(lo + k);
        b = map_translate(c3[0], c3[1], c3[2], a.tab64);
        if (!b) break;
      } else {
        b = oriented(r.ali_from - e - 1);
      }
      if (low_freq(a, h, r.qstart - e, b)) break;
      bp[-e] = b;
      if (fr >= 0) { cp[-3 * e] = c3[0]; cp[-3 * e + 1] = c3[1]; cp[-3 * e + 2] = c3[2]; }
      es = e;
    }
  }
  if (et == 0) {
    for (int e = 1; e <= a.extend_pos && r.qend + e <= a.ncols_ref[h]; e++) {
      char b, c3[3] = {0, 0, 0};
      if (fr >= 0) {
        const int hi = 3 * r.ali_to + fr + 3 * e;
        if (hi > rlen) break;
        for (int k = 0; k < 3; k++) c3[k] = oriented(hi - 3 + k);
        b = map_translate(c3[0], c3[1], c3[2], a.tab64);
        if (!b) break;
      } else {
        if (r.ali_to + e > rlen) break;
        b = oriented(r.ali_to + e - 1);
      }
      if (low_freq(a, h, r.qend + e, b)) break;
      bp[col + e - 1] = b;
      if (fr >= 0) { cp[3 * (col + e - 1)] = c3[0]; cp[3 * (col + e - 1) + 1] = c3[1]; cp[3 * (col + e - 1) + 2] = c3[2]; }
      ee = e;
    }
  }
  {
    const int w = fr >= 0 ? 3 : 1;
    if (!rev) { r.read_from -= w * es; r.read_to += w * ee; }
    else { r.read_to += w * es; r.read_from -= w * ee; }
  }
  r.qstart -= es; r.qend += ee;
  r.ext_start = es; r.ext_end = ee;
  r.col_off = (int)a.col_off[h] + a.extend_pos - es;
  r.ncols = col + es + ee;
  r.n_unmap = nun;
  r.keep = 1;
  a.rec[h] = r;
}

// ---- read_hit.consensus (lib/assemble.py:309-633): histogram accumulation ----
struct AccumArgs {
  int nhits;
  const pa_maprec *rec;
  const char *base_pool, *codon_pool;
  const int *frame, *bin, *rep;
  const long long *bin_row;               // per bin: first histogram row (row = reference column - 1)
  int *hist;                              // [rows][3][32] counts
  int *first;                             // [rows][3][32] smallest contributing hit index (INT_MAX = none)
  int *qrange;                            // [nbins][2]: min qstart, max qend
};

__global__ void accum_kernel(AccumArgs a) {
  const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (h >= a.nhits) return;
  const pa_maprec r = a.rec[h];
  const int b = a.bin[h];
  if (!r.keep || b < 0) return;
  const int rep = a.rep[h];
  const bool codon = a.frame[h] >= 0;
  for (int c = lane; c < r.ncols; c += 32) {
    const size_t row = (size_t)a.bin_row[b] + (r.qstart + c - 1);
    for (int k = 0; k < (codon ? 3 : 1); k++) {
      const char ch = codon ? a.codon_pool[3 * ((size_t)r.col_off + c) + k] : a.base_pool[(size_t)r.col_off + c];
      const size_t cell = (row * 3 + k) * 32 + comp_sym(ch);
      atomicAdd(a.hist + cell, rep);
      atomicMin(a.first + cell, h);
    }
  }
  if (lane == 0) {
    atomicMin(a.qrange + 2 * b, r.qstart);
    atomicMax(a.qrange + 2 * b + 1, r.qend);
  }
}

// ---- merge_hits (lib/assemble.py:889-954), histogram part: a segmented left fold ----
// Chain j folds items chain_off[j] .. chain_off[j+1]-1 in order (newhit = merge_hits(newhit, hit) repeatedly).  Item i holds the
// counts of its reference columns item_q0[i] .. item_q0[i] + item_ncols[i] - 1 as [column][3][32] at row item_row[i], with
// key_in (same shape) = the insertion rank of every symbol inside the item's own dictionaries.  Per chain the kernel writes the
// summed counts over columns 1 .. chain_ncols[j] plus, for the dict insertion order of the result, the first item that brought
// each symbol (position inside the chain) and that item's own rank for it.
struct MergeArgs {
  int nchains;
  const int *chain_off, *item_q0, *item_ncols;
  const long long *item_row, *chain_row;
  const int *chain_ncols;
  const int *hist_in, *key_in;
  int *hist_out, *who_out, *key_out;
};

__global__ void merge_kernel(MergeArgs a) {
  const int j = blockIdx.y;
  const int cells = a.chain_ncols[j] * 96;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
    const int col = c / 96 + 1, rest = c % 96;
    int sum = 0, who = INT_MAX, key = INT_MAX;
    for (int i = a.chain_off[j]; i < a.chain_off[j + 1]; i++) {
      const int rel = col - a.item_q0[i];
      if (rel < 0 || rel >= a.item_ncols[i]) continue;
      const size_t cell = ((size_t)a.item_row[i] + rel) * 96 + rest;
      const int v = a.hist_in[cell];
      if (v != 0 && who == INT_MAX) { who = i - a.chain_off[j]; key = a.key_in[cell]; }
      sum += v;
    }
    const size_t o = (size_t)a.chain_row[j] * 96 + c;
    a.hist_out[o] = sum;
    a.who_out[o] = who;
    a.key_out[o] = key;
  }
}

}  // namespace pa
