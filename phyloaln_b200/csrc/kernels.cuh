// kernels.cuh -- sm_100a kernels of the PhyloAln search hot path (filters).
//
//  translate_kernel   six-frame / codon translation or reverse complement + digitisation
//                     (reference lib/aln.py:233-259; HBM-bound)
//  msv_kernel<Q>      SSV pass in DPX s16x2 lanes (VIADDMNMX.S16x2.RELU + VIMNMX3.S16x2) with the
//                     striped emission table in shared memory; pairs that could use the J state
//                     or reach the F1 threshold are re-run inline with the exact unsigned-byte
//                     MSV recurrence
//  bias_kernel        2-state composition-bias filter (A.6)
//  vit_kernel         Viterbi filter in exact signed-16 saturating semantics evaluated in s32 DPX
//                     (VIADDMNMX with the -32768 clamp fused), parallel max-plus scan for D->D (A.5)
//  fwd_kernel         Forward filter in scaled probability space, canonical 32-lane blocked
//                     evaluation order (DESIGN.md "Canonical float order")
//
// All DP kernels: one warp per (profile, target) pair.  Lane l owns the contiguous model block
// k = l*B+1 .. (l+1)*B, B = ceil(M/32); per-profile tables are stored [j][lane] so that every
// table access of a warp is one fully coalesced 128/512-byte request.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pa {

#define PA_FULL 0xffffffffu
#define PA_LN2 0.69314718055994529
#define PA_MSV_DEAD (-20000)
#define PA_NEG16 (-32768)
#define PA_NEGBIG (-(1 << 29))
#define PA_RESCALE 8192.0f

struct DevHmm {
  int M, Q, B, abc;             // Q: MSV words per lane; B = ceil(M/32)
  int Bv, rank;                 // Viterbi lane-block size (class-rounded); position in (Bv, index) order
  int nrows;                    // Kp
  unsigned long long msv_off;   // uint32 words into msv pool
  unsigned long long msvh_off;  // uint32 words into the fp16x2 pool of the TMEM kernel ([x][q][lane])
  unsigned long long vit_off;   // ints into vit pool: trA(int4) | trB(int4) | ps | em[Kp]
  unsigned long long fwd_off;   // floats into fwd pool: trA | trB | bwA | bwB | em[Kp]
  unsigned long long v16_off;   // uint32 words into the packed Viterbi pool of vit16_kernel (vit16.cuh)
  int BH, v16_guard;            // words per lane of vit16_kernel (class-rounded ceil(M/64)); headroom its values need below 32767
  int nT;                       // wide profiles (M > 2047, wide.cuh): model tiles of msv_wide_kernel (0: not wide); msvh_off -> nT tile tables,
  int NWv, NW16;                // msv_off -> plain [x][q][lane] table; warps per pair of vit_wide_kernel / vit16_wide_kernel (1: not wide)
  int tbm, tec, bias, base;
  int base_w, xw_E;
  float scale_b, scale_w;
  float thrF1, thrF2m, thrF2v, thrF3;  // bit-score thresholds equivalent to P <= F
  float t00, t01, t10, t11;            // bias-filter transitions
  float eo1[32];
  float ev[6];
};

struct Pass1 { uint32_t t; uint16_t h; int16_t xJ; };
struct Pass2 { uint32_t t; uint16_t h; uint16_t flags; float filtersc; };  // flags bit0: needs Viterbi
struct Pass4 { uint32_t t; uint16_t h; uint16_t pad; float filtersc; int fexp; float fmant; };

struct Targets {
  const uint8_t *dsq;        // digital residues
  const uint32_t *off;       // byte offset of each target (4-byte aligned)
  const uint32_t *len;       // residues
  const uint16_t *lidx;      // index into the distinct-length tables
  uint32_t n;
};

struct LenTabs {             // per distinct target length
  const int *L;
  const float *nullsc;
  const uint8_t *tjb;        // MSV J->B byte cost
  const int16_t *xw_move;    // Viterbi N->B / J->B / C->T word
  const float *biasA;        // (float)L*logf(p1)        (bias-filter length terms, host libm)
  const float *biasB;        // logf(1-p1)
  int n;
};

// ---------------------------------------------------------------------------------------------
// translation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nt_mask(uint8_t c) {  // bit: T=1 C=2 A=4 G=8; 0 invalid
  switch (c & 0xDF) {                                  // upper-case
    case 'T': case 'U': return 1;
    case 'C': return 2;
    case 'A': return 4;
    case 'G': return 8;
    case 'R': return 12;
    case 'Y': return 3;
    case 'M': return 6;
    case 'K': return 9;
    case 'S': return 10;
    case 'W': return 5;
    case 'H': return 7;
    case 'B': return 11;
    case 'V': return 14;
    case 'D': return 13;
    case 'N': return 15;
    default: return 0;
  }
}
__device__ __forceinline__ int comp_mask(int m) { return ((m & 1) << 2) | ((m & 4) >> 2) | ((m & 2) << 2) | ((m & 8) >> 2); }

// digital amino codes: A0 C1 D2 E3 F4 G5 H6 I7 K8 L9 M10 N11 P12 Q13 R14 S15 T16 V17 W18 Y19 -20 B21 J22 Z23 O24 U25 X26 *27 ~28
__device__ __forceinline__ int codon_code(int m0, int m1, int m2, const uint8_t *tab64) {
  if (__popc(m0) == 1 && __popc(m1) == 1 && __popc(m2) == 1)
    return tab64[(__ffs(m0) - 1) * 16 + (__ffs(m1) - 1) * 4 + (__ffs(m2) - 1)];
  // ambiguous codon (Biopython rules)
  uint32_t seen = 0;
  int nstop = 0, ntot = 0;
  for (int a = 0; a < 4; a++)
    if (m0 >> a & 1)
      for (int b = 0; b < 4; b++)
        if (m1 >> b & 1)
          for (int c = 0; c < 4; c++)
            if (m2 >> c & 1) {
              int r = tab64[a * 16 + b * 4 + c];
              ntot++;
              if (r == 27) nstop++;
              else seen |= 1u << r;
            }
  if (nstop == ntot) return 27;
  if (nstop > 0) return 26;
  if (__popc(seen) == 1) return __ffs(seen) - 1;
  if ((seen & ~((1u << 2) | (1u << 11))) == 0) return 21;  // D,N -> B
  if ((seen & ~((1u << 3) | (1u << 13))) == 0) return 23;  // E,Q -> Z
  if ((seen & ~((1u << 7) | (1u << 9))) == 0) return 22;   // I,L -> J
  return 26;
}
// DNA digital codes "ACGT-RYMKSWHBVDN*~" from a T1 C2 A4 G8 mask
__device__ __forceinline__ int dna_code(int m) {
  const uint8_t tab[16] = {15, 3, 1, 6, 0, 10, 7, 11, 2, 8, 9, 12, 5, 14, 13, 15};
  return tab[m];
}

__device__ __forceinline__ uint32_t target_base(unsigned long long ntoff, uint32_t r) {
  return (uint32_t)(((2ull * ntoff + 3ull) & ~3ull) + 32ull * r);
}

// Generic per-read routine (any length, both molecule types): every lane converts the nucleotides it
// needs straight from global memory.  Used for nucleotide targets and for reads longer than the
// staging buffer of translate_kernel.
#define PA_TR_MAXL 1024
static __device__ void translate_read_generic(const char *__restrict__ nt, unsigned long long o, int L, uint32_t r, int moltype, int nfwd,
                                       int nrev, const uint8_t *tab64, uint8_t *__restrict__ out, uint32_t *__restrict__ tgt_off,
                                       uint32_t *__restrict__ tgt_len, int *__restrict__ err, int lane) {
  const int nframes = nfwd + nrev;
  auto mask_at = [&](int i) -> int {
    int m = nt_mask((uint8_t)nt[o + i]);
    if (m == 0) { atomicExch(err, 1); m = 15; }
    return m;
  };
  // complement of the letter at i as a mask.  'U' is not in Biopython's DNA complement table (Bio.Seq >= 1.80), so
  // reverse_complement() leaves it in place and it is read as T on the reverse strand too (oracle: orc_revcomp)
  auto comp_at = [&](int i) -> int {
    const uint8_t c = (uint8_t)nt[o + i] & 0xDF;
    return c == 'U' ? 1 : comp_mask(mask_at(i));
  };
  uint32_t pos = target_base(o, r);
  for (int f = 0; f < nframes; f++) {
    const bool rev = (moltype == 1) ? (f == 1) : (f >= nfwd);
    const int j = (moltype == 1) ? 0 : (rev ? f - nfwd : f);  // frame offset 0..2
    const int step = moltype == 1 ? 1 : 3;
    const int nres = moltype == 1 ? L : ((L - j) >= 3 ? (L - j) / 3 : 0);
    const int nwords = (nres + 3) >> 2;
    for (int w = lane; w < nwords; w += 32) {
      uint32_t word = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int c = 4 * w + q;
        int code = 0;
        if (c < nres) {
          const int p0 = j + step * c;
          if (moltype == 1) {
            int m = rev ? comp_at(L - 1 - p0) : mask_at(p0);
            code = dna_code(m);
          } else {
            int m0, m1, m2;
            if (!rev) { m0 = mask_at(p0); m1 = mask_at(p0 + 1); m2 = mask_at(p0 + 2); }
            else { m0 = comp_at(L - 1 - p0); m1 = comp_at(L - 2 - p0); m2 = comp_at(L - 3 - p0); }
            code = codon_code(m0, m1, m2, tab64);
          }
        }
        word |= (uint32_t)code << (8 * q);
      }
      *reinterpret_cast<uint32_t *>(out + pos + 4 * w) = word;
    }
    if (lane == 0) { tgt_off[(size_t)r * nframes + f] = pos; tgt_len[(size_t)r * nframes + f] = nres; }
    pos += (nres + 3) & ~3;
  }
}

static __global__ void __launch_bounds__(256) translate_generic_kernel(const char *__restrict__ nt, const unsigned long long *__restrict__ off,
                                                                uint32_t nreads, int moltype, int nfwd, int nrev,
                                                                const uint8_t *__restrict__ tab64, uint8_t *__restrict__ out,
                                                                uint32_t *__restrict__ tgt_off, uint32_t *__restrict__ tgt_len,
                                                                int *__restrict__ err) {
  __shared__ uint8_t s_tab[64];
  if (threadIdx.x < 64) s_tab[threadIdx.x] = tab64[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (uint32_t r = blockIdx.x * wpb + wib; r < nreads; r += gridDim.x * wpb) {
    const unsigned long long o = off[r];
    translate_read_generic(nt, o, (int)(off[r + 1] - o), r, moltype, nfwd, nrev, s_tab, out, tgt_off, tgt_len, err, lane);
  }
}

// Exact IUPAC path for one triplet (Biopython rules): forward residue and the residue of the reverse
// complement.  Kept out of line: it runs for the ~0.3 % of triplets that contain a non-ACGTU letter.
// t0..t2: the s_nt bytes of the three letters (bits 4-7 IUPAC mask; bit 0 set for 'U', whose complement is itself = T)
static __device__ __noinline__ int translate_ambiguous(int t0, int t1, int t2, const uint8_t *tab64) {
  const int m0 = t0 >> 4, m1 = t1 >> 4, m2 = t2 >> 4;
  // 'U' = exact-path flag (bit 2) + marker (bit 0); bit 0 alone is part of the 2-bit code of a plain letter
  const int r0 = (t0 & 5) == 5 ? 1 : comp_mask(m0), r1 = (t1 & 5) == 5 ? 1 : comp_mask(m1), r2 = (t2 & 5) == 5 ? 1 : comp_mask(m2);
  auto one = [&](int a0, int a1, int a2) -> int {
    uint32_t seen = 0;
    int nstop = 0, ntot = 0;
#pragma unroll 1
    for (int x = a0; x; x &= x - 1)
#pragma unroll 1
      for (int y = a1; y; y &= y - 1)
#pragma unroll 1
        for (int z = a2; z; z &= z - 1) {
          const int r = tab64[(__ffs(x) - 1) * 16 + (__ffs(y) - 1) * 4 + (__ffs(z) - 1)];
          ntot++;
          if (r == 27) nstop++;
          else seen |= 1u << r;
        }
    if (nstop == ntot) return 27;
    if (nstop > 0) return 26;
    if (__popc(seen) == 1) return __ffs(seen) - 1;
    if ((seen & ~((1u << 2) | (1u << 11))) == 0) return 21;  // D,N -> B
    if ((seen & ~((1u << 3) | (1u << 13))) == 0) return 23;  // E,Q -> Z
    if ((seen & ~((1u << 7) | (1u << 9))) == 0) return 22;   // I,L -> J
    return 26;
  };
  return one(m0, m1, m2) | (one(r2, r1, r0) << 8);
}

// Translation (reference lib/aln.py:233-259), HBM-bound design.  One warp per read:
//   1. coalesced byte loads; a 256-byte table turns each character into a 2-bit code (A0 C1 T/U2 G3,
//      either case; complement = code ^ 2), an "ambiguous" flag and its 4-bit IUPAC mask;
//   2. lane q forms the 6-bit index of the triplet starting at q from its own code and its two right
//      neighbours (warp shuffles, no shared-memory traffic) and looks it up twice in 64-byte tables:
//      the forward codon table and the table of the reverse-complemented codon.  Position q yields
//      exactly one forward residue (frame q % 3) and one reverse residue (frame (L-3-q) % 3);
//   3. the two residues are scattered as bytes into the read's frame-major image in shared memory, which
//      is then copied out with coalesced 32-bit stores.
// HBM traffic per read: L bytes in, six 4-byte-padded frames out (446 B algorithmic for 150 bp).
template <int NFWD, int NREV>
static __global__ void __launch_bounds__(256) translate_kernel(const char *__restrict__ nt, const unsigned long long *__restrict__ off,
                                                        uint32_t nreads, const uint8_t *__restrict__ tab64,
                                                        uint8_t *__restrict__ out, uint32_t *__restrict__ tgt_off,
                                                        uint32_t *__restrict__ tgt_len, int *__restrict__ err) {
  constexpr int NFR = NFWD + NREV;
  __shared__ uint8_t s_tab[64], s_tabF[64], s_tabR[64], s_nt[256];
  __shared__ __align__(16) uint8_t s_out[8][2 * PA_TR_MAXL + 32];
  if (threadIdx.x < 64) {
    s_tab[threadIdx.x] = tab64[threadIdx.x];
    // my code order A0 C1 T2 G3 -> tab64 digit order T0 C1 A2 G3
    const int f = threadIdx.x, c0 = f >> 4, c1 = (f >> 2) & 3, c2 = f & 3;
    auto m = [](int c) { return c == 0 ? 2 : c == 2 ? 0 : c; };
    s_tabF[f] = tab64[m(c0) * 16 + m(c1) * 4 + m(c2)];
    s_tabR[f] = tab64[m(c2 ^ 2) * 16 + m(c1 ^ 2) * 4 + m(c0 ^ 2)];
  }
  {  // per character: bits 0-1 code (A/C/G/T only), bit 2 = not A/C/G/T (exact path), bit 3 = not a nucleotide letter at all,
     // bits 4-7 IUPAC mask; 'U' takes the exact path with bit 0 as its marker (forward: T; complement: itself, see above)
    const int c = threadIdx.x;
    const int mk = nt_mask((uint8_t)c);
    const bool isU = (c & 0xDF) == 'U';
    const bool plain = (mk == 1 || mk == 2 || mk == 4 || mk == 8) && !isU;
    s_nt[c] = (uint8_t)((plain ? ((c >> 1) & 3) : (isU ? 1 : 0)) | (plain ? 0 : 4) | (mk == 0 ? 8 : 0) | ((mk ? mk : 15) << 4));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = __shfl_sync(PA_FULL, (int)(threadIdx.x >> 5), 0), wpb = blockDim.x >> 5;
  uint8_t *so = s_out[wib];
  for (uint32_t r = blockIdx.x * wpb + wib; r < nreads; r += gridDim.x * wpb) {
    const unsigned long long o = off[r];
    const int L = (int)(off[r + 1] - o);
    if (L > PA_TR_MAXL) {
      translate_read_generic(nt, o, L, r, 0, NFWD, NREV, s_tab, out, tgt_off, tgt_len, err, lane);
      continue;
    }
    // frame geometry (warp-uniform): frame offset j has (L-j)/3 residues, frames are 4-byte padded
    const int n0 = L >= 3 ? L / 3 : 0, n1 = L >= 4 ? (L - 1) / 3 : 0, n2 = L >= 5 ? (L - 2) / 3 : 0;
    const int p0 = (n0 + 3) & ~3, p1 = (n1 + 3) & ~3, p2 = (n2 + 3) & ~3;
    const int fb1 = p0, fb2 = p0 + p1;
    const int rb0 = NFWD == 3 ? p0 + p1 + p2 : p0, rb1 = rb0 + p0, rb2 = rb1 + p1;
    const int tot = rb0 + (NREV == 3 ? p0 + p1 + p2 : 0);
    const uint32_t pos = target_base(o, r);
    __syncwarp();
    // the last word of every frame carries the zero padding
    if (lane < NFR) {
      const int j = lane < NFWD ? lane : lane - NFWD;
      const int b = lane < NFWD ? (j == 0 ? 0 : j == 1 ? fb1 : fb2) : (j == 0 ? rb0 : j == 1 ? rb1 : rb2);
      const int pj = j == 0 ? p0 : j == 1 ? p1 : p2;
      if (pj) *reinterpret_cast<uint32_t *>(so + b + pj - 4) = 0u;
    }
    __syncwarp();
    const int npos = L - 2;  // triplet start positions
    int cur = lane < L ? (int)s_nt[(uint8_t)nt[o + lane]] : 4;
    for (int q0 = 0; q0 < npos; q0 += 32) {
      const int i1 = q0 + 32 + lane;
      const int nxt = i1 < L ? (int)s_nt[(uint8_t)nt[o + i1]] : 4;
      const int a1 = __shfl_sync(PA_FULL, cur, (lane + 1) & 31), b1 = __shfl_sync(PA_FULL, nxt, (lane + 1) & 31);
      const int a2 = __shfl_sync(PA_FULL, cur, (lane + 2) & 31), b2 = __shfl_sync(PA_FULL, nxt, (lane + 2) & 31);
      const int c1 = lane < 31 ? a1 : b1, c2 = lane < 30 ? a2 : b2;
      const int q = q0 + lane;
      if (q < npos) {
        int aaF, aaR;
        if (((cur | c1 | c2) & 4) == 0) {
          const int f = (cur & 3) * 16 + (c1 & 3) * 4 + (c2 & 3);
          aaF = s_tabF[f];
          aaR = s_tabR[f];
        } else {
          if ((cur | c1 | c2) & 8) atomicExch(err, 1);
          const int both = translate_ambiguous(cur, c1, c2, s_tab);
          aaF = both & 0xff;
          aaR = both >> 8;
        }
        const int cf = (q * 43691) >> 17, jf = q - 3 * cf;  // q / 3, q % 3 (q < 32768)
        if (jf < NFWD) so[(jf == 0 ? 0 : jf == 1 ? fb1 : fb2) + cf] = (uint8_t)aaF;
        if (NREV) {
          const int qr = npos - 1 - q;
          const int cr = (qr * 43691) >> 17, jr = qr - 3 * cr;
          so[(jr == 0 ? rb0 : jr == 1 ? rb1 : rb2) + cr] = (uint8_t)aaR;
        }
      }
      cur = nxt;
    }
    __syncwarp();
    const uint32_t *so4 = reinterpret_cast<const uint32_t *>(so);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + pos);
    for (int w = lane; w < (tot >> 2); w += 32) dst[w] = so4[w];
    if (lane < NFR) {
      const int j = lane < NFWD ? lane : lane - NFWD;
      const int b = lane < NFWD ? (j == 0 ? 0 : j == 1 ? fb1 : fb2) : (j == 0 ? rb0 : j == 1 ? rb1 : rb2);
      tgt_off[(size_t)r * NFR + lane] = pos + b;
      tgt_len[(size_t)r * NFR + lane] = (j == 0 ? n0 : j == 1 ? n1 : n2);
    }
  }
}

static __global__ void lidx_kernel(const uint32_t *__restrict__ len, const uint16_t *__restrict__ L2li, uint16_t *__restrict__ lidx, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lidx[i] = L2li[len[i]];
}

// protein / pre-digitised targets: ASCII -> digital codes
static __global__ void digitize_kernel(const char *__restrict__ in, const unsigned long long *__restrict__ in_off,
                                const uint32_t *__restrict__ tgt_off, uint32_t n, int abc, uint8_t *__restrict__ out) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (uint32_t s = blockIdx.x * wpb + wib; s < n; s += gridDim.x * wpb) {
    unsigned long long o = in_off[s];
    int L = (int)(in_off[s + 1] - o);
    for (int i = lane; i < L; i += 32) {
      uint8_t c = (uint8_t)in[o + i] & 0xDF;
      int code;
      if (abc == 0) {
        const char *syms = "ACDEFGHIKLMNPQRSTVWY-BJZOUX";
        code = 26;
        if ((uint8_t)in[o + i] == '*') code = 27;
        else
          for (int q = 0; q < 27; q++)
            if (syms[q] == (char)c) { code = q; break; }
      } else {
        int m = nt_mask(c);
        code = dna_code(m == 0 ? 15 : m);
      }
      out[tgt_off[s] + i] = (uint8_t)code;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SSV / MSV filter
// ---------------------------------------------------------------------------------------------
template <int Q>
__device__ __forceinline__ void msv_load_row(const uint32_t *row, int lane, uint32_t (&d)[Q]) {
  constexpr int N4 = Q / 4, R2 = (Q % 4) >= 2 ? 1 : 0, R1 = Q % 2;
#pragma unroll
  for (int g = 0; g < N4; g++) {
    uint4 v = *reinterpret_cast<const uint4 *>(row + g * 128 + lane * 4);
    d[4 * g] = v.x; d[4 * g + 1] = v.y; d[4 * g + 2] = v.z; d[4 * g + 3] = v.w;
  }
  if (R2) {
    uint2 v = *reinterpret_cast<const uint2 *>(row + N4 * 128 + lane * 2);
    d[4 * N4] = v.x; d[4 * N4 + 1] = v.y;
  }
  if (R1) d[Q - 1] = row[N4 * 128 + R2 * 64 + lane];
}

__device__ __forceinline__ int warp_max16x2(uint32_t m) {
  int v = max((int)(short)(m & 0xffff), (int)(short)(m >> 16));
  return __reduce_max_sync(PA_FULL, v);
}

struct MsvArgs {
  const DevHmm *hmms;
  const int *hmm_list;     // profiles with this Q
  int nlist;
  const uint32_t *tab;     // msv table pool (s16x2 striped for msv_kernel, fp16x2 [x][q][lane] for msv_tmem_kernel)
  Targets tg;
  const uint8_t *tjb_li;   // per length index
  const int16_t *thrJ;     // [hmm*nL + li] minimal passing xJ (0..256)
  const uint16_t *scJ;     // [hmm*nL + li] fp16 bits of the sticky-SSV scale 32*ceil(2048/Cp) (Cp = 1: +inf; Cp <= 0: 0xffff)
  int nL;
  uint32_t tile;           // targets per work item
  uint32_t super_tiles;    // tiles per L2-sized super-tile: all profiles run over one super-tile of targets before the next
  unsigned long long *work;  // msv_tmem_kernel: next work item (dynamic distribution over CTAs), zeroed per launch
  Pass1 *out;
  unsigned long long *out_count;
  unsigned long long out_cap;
  unsigned long long *n_ssv_cand;
  int all_pairs;           // 1: emit every pair with its exact MSV result (score_all)
  uint32_t zero;           // always 0; a run-time value so ptxas cannot re-materialise {0,0} per use
};

// Work item -> (profile slot, target tile).  Items are profile-major INSIDE a super-tile of `T` target tiles (neighbouring items
// share the profile, so its table is reloaded rarely) and super-tile-major overall, so that a chunk whose targets exceed the L2 is
// streamed from DRAM once per super-tile instead of once per profile.  T >= ntiles gives the plain profile-major order.
__host__ __device__ __forceinline__ void msv_item_decode(unsigned long long item, uint32_t nlist, uint32_t ntiles, uint32_t T, uint32_t &slot,
                                                uint32_t &tile) {
  const uint32_t full = ntiles / T;
  const unsigned long long per = (unsigned long long)nlist * T, items_full = per * full;
  if (item < items_full) {
    const uint32_t s = (uint32_t)(item / per), r = (uint32_t)(item % per);
    slot = r / T;
    tile = s * T + r % T;
  } else {
    const uint32_t r = (uint32_t)(item - items_full), Ts = ntiles - full * T;
    slot = r / Ts;
    tile = full * T + r % Ts;
  }
}

template <int Q>
static __global__ void __launch_bounds__(256, 2) msv_kernel(MsvArgs a) {
  extern __shared__ __align__(16) uint32_t s_tab[];  // nrows * Q*32 words
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint32_t ntiles = (a.tg.n + a.tile - 1) / a.tile;
  const unsigned long long nitems = (unsigned long long)a.nlist * ntiles;
  int cur_h = -1;
  unsigned long long ncand = 0;
  constexpr int ROWW = Q * 32;
  // Q = M/64 + 1, so the very last striped cell (lane 31, high half, word Q-1) is always padding and
  // stays 0: the row carry can wrap around from lane 31 to lane 0 without a select.
  const int lanem1 = (lane + 31) & 31;
  const uint32_t zero = a.zero;
  for (unsigned long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    uint32_t slot, tile;
    msv_item_decode(item, (uint32_t)a.nlist, ntiles, a.super_tiles, slot, tile);
    const int h = a.hmm_list[slot];
    const DevHmm &hm = a.hmms[h];
    if (h != cur_h) {
      __syncthreads();
      const uint4 *src = reinterpret_cast<const uint4 *>(a.tab + hm.msv_off);
      uint4 *dst = reinterpret_cast<uint4 *>(s_tab);
      const int n4 = hm.nrows * ROWW / 4;
      for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
      __syncthreads();
      cur_h = h;
    }
    const int bias = hm.bias, base = hm.base, tec = hm.tec, tbm = hm.tbm;
    const uint32_t t1 = min(a.tg.n, (tile + 1) * a.tile);
    for (uint32_t t = tile * a.tile + wib; t < t1; t += wpb) {
      const int L = (int)a.tg.len[t];
      if (L == 0) continue;
      const uint8_t *p = a.tg.dsq + a.tg.off[t];
      const int li = a.tg.lidx[t];
      const int tjb = a.tjb_li[li];
      const int T = a.thrJ[(size_t)h * a.nL + li];
      const int tjbm = (tjb + tbm) & 0xff;
      const int xB0 = max(base - tjbm, 0);
      int C = min(255 - bias, min(base + 1 + tec, T + tec));
      const int Cp = a.all_pairs ? 0 : C - xB0;
      // ---- SSV pass: u(i,k) = relu(u(i-1,k-1) + (bias - rbv)), running max
      uint32_t w[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) w[q] = 0;
      uint32_t m = 0;
      if (Cp > 0) {
        // each lane fetches one residue of a 32-row block and turns it into a table-row offset;
        // rows then get their offset by shuffle: no ALU work for addressing in the row body
        uint32_t myoff = (lane < L) ? (uint32_t)p[lane] * ROWW : 0u;
        for (int i0 = 0; i0 < L; i0 += 32) {
          const uint32_t curoff = myoff;
          if (i0 + 32 < L) myoff = (i0 + 32 + lane < L) ? (uint32_t)p[i0 + 32 + lane] * ROWW : 0u;
          const int nr = min(32, L - i0);
#define PA_SSV_ROW(rr)                                                                     \
  {                                                                                        \
    const uint32_t ro = __shfl_sync(PA_FULL, curoff, (rr));                                \
    uint32_t d[Q];                                                                         \
    msv_load_row<Q>(s_tab + ro, lane, d);                                                  \
    const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);                            \
    const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);                              \
    _Pragma("unroll") for (int q = Q - 1; q >= 1; q--) w[q] = __viaddmax_s16x2(w[q - 1], d[q], zero); \
    w[0] = __viaddmax_s16x2(carry, d[0], zero);                                            \
    _Pragma("unroll") for (int q = 0; q + 1 < Q; q += 2) m = __vimax3_s16x2(m, w[q], w[q + 1]); \
    if (Q & 1) m = __vmaxs2(m, w[Q - 1]);                                                  \
  }
          int r = 0;
          for (; r + 4 <= nr; r += 4) {
            PA_SSV_ROW(r) PA_SSV_ROW(r + 1) PA_SSV_ROW(r + 2) PA_SSV_ROW(r + 3)
          }
          for (; r < nr; r++) PA_SSV_ROW(r)
#undef PA_SSV_ROW
        }
      }
      const int U = (Cp > 0) ? warp_max16x2(m) : 0;
      if (U < Cp) continue;
      ncand++;
      // ---- exact MSV recurrence (J state, overflow) for candidates
      int xJ = 0, xB = xB0;
      bool overflow = false;
#pragma unroll
      for (int q = 0; q < Q; q++) w[q] = 0;
      for (int i = 0; i < L && !overflow; i++) {
        const int x = p[i];
        uint32_t d[Q];
        msv_load_row<Q>(s_tab + x * ROWW, lane, d);
        const uint32_t xBv = (uint32_t)xB * 0x10001u;
        const uint32_t tt = __shfl_sync(PA_FULL, w[Q - 1], lanem1);
        const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
        uint32_t rm = 0;
#pragma unroll
        for (int q = Q - 1; q >= 1; q--) {
          w[q] = __viaddmax_s16x2(__vmaxs2(w[q - 1], xBv), d[q], zero);
          rm = __vmaxs2(rm, w[q]);
        }
        w[0] = __viaddmax_s16x2(__vmaxs2(carry, xBv), d[0], zero);
        rm = __vmaxs2(rm, w[0]);
        int xE = warp_max16x2(rm);
        if (xE >= 255 - bias) { overflow = true; break; }
        xE = max(xE - tec, 0);
        xJ = max(xJ, xE);
        xB = max(max(base, xJ) - tjbm, 0);
      }
      if (overflow) xJ = -1;
      if (a.all_pairs || overflow || xJ >= T) {
        if (lane == 0) {
          unsigned long long idx = atomicAdd(a.out_count, 1ull);
          if (idx < a.out_cap) {
            Pass1 r;
            r.t = t; r.h = (uint16_t)h; r.xJ = (int16_t)xJ;
            a.out[idx] = r;
          }
        }
      }
    }
  }
  if (lane == 0 && ncand) atomicAdd(a.n_ssv_cand, ncand);
}

// per (profile, length) minimal passing MSV xJ byte
static __global__ void msv_threshold_kernel(const DevHmm *hmms, int nh, LenTabs lt, int16_t *thrJ, uint16_t *scJ, int all_pass) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nh * lt.n) return;
  int h = idx / lt.n, li = idx % lt.n;
  const DevHmm &hm = hmms[h];
  int tjb = lt.tjb[li];
  float nullsc = lt.nullsc[li];
  int T = 256;
  for (int xJ = 0; xJ < 256; xJ++) {
    float usc = ((float)(xJ - tjb) - (float)hm.base);
    usc /= hm.scale_b;
    usc -= 3.0f;
    float seq = (float)((double)(usc - nullsc) / PA_LN2);
    if (seq >= hm.thrF1) { T = xJ; break; }
  }
  if (all_pass) T = 0;
  thrJ[idx] = (int16_t)T;
  // candidate threshold of the SSV pass (same expression as in the MSV kernels) -> sticky scale
  const int tjbm = (tjb + hm.tbm) & 0xff;
  const int xB0 = max(hm.base - tjbm, 0);
  const int C = min(255 - hm.bias, min(hm.base + 1 + hm.tec, T + hm.tec));
  const int Cp = C - xB0;
  uint16_t sc;
  if (Cp <= 0) sc = 0xffffu;
  else if (Cp == 1) sc = 0x7c00u;
  else sc = __half_as_ushort(__int2half_rn(32 * ((2048 + Cp - 1) / Cp)));
  scJ[idx] = sc;
}

__device__ __forceinline__ float msv_score_from_xJ(const DevHmm &hm, int xJ, int tjb) {
  if (xJ < 0) return __int_as_float(0x7f800000);
  float usc = ((float)(xJ - tjb) - (float)hm.base);
  usc /= hm.scale_b;
  usc -= 3.0f;
  return usc;
}

// ---------------------------------------------------------------------------------------------
// bias filter: one thread per MSV survivor
// ---------------------------------------------------------------------------------------------
static __device__ float bias_filtersc(const DevHmm &hm, const uint8_t *p, int L, float lenA, float lenB) {
  float d0 = 0.999f, d1 = 0.001f * hm.eo1[p[0]];
  int sexp = 0;
  for (int i = 1;; i++) {
    float mx = fmaxf(d0, d1);
    int e = ilogbf(mx);
    d0 = scalbnf(d0, -e);
    d1 = scalbnf(d1, -e);
    sexp += e;
    if (i == L) break;
    float n0 = __fadd_rn(__fmul_rn(d0, hm.t00), __fmul_rn(d1, hm.t10));
    float n1 = __fmul_rn(__fadd_rn(__fmul_rn(d0, hm.t01), __fmul_rn(d1, hm.t11)), hm.eo1[p[i]]);
    d0 = n0;
    d1 = n1;
  }
  float tot = __fadd_rn(d0, d1);
  float hmmsc = (float)(log((double)tot) + (double)sexp * PA_LN2);
  return __fadd_rn(__fadd_rn(hmmsc, lenA), lenB);
}

struct BiasArgs {
  const DevHmm *hmms;
  Targets tg;
  LenTabs lt;
  const Pass1 *in;
  unsigned long long n_in;
  Pass2 *out;
  unsigned long long *out_count;
  unsigned long long out_cap;
  int do_bias;
  float *bias_out;  // optional (score_all): filtersc per input record
};

static __global__ void bias_kernel(BiasArgs a) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= a.n_in) return;
  Pass1 r = a.in[i];
  const DevHmm &hm = a.hmms[r.h];
  const int L = (int)a.tg.len[r.t];
  const int li = a.tg.lidx[r.t];
  const float nullsc = a.lt.nullsc[li];
  const float usc = msv_score_from_xJ(hm, r.xJ, a.lt.tjb[li]);
  float filtersc = nullsc;
  if (a.do_bias && L > 0) filtersc = bias_filtersc(hm, a.tg.dsq + a.tg.off[r.t], L, a.lt.biasA[li], a.lt.biasB[li]);
  if (a.bias_out) a.bias_out[i] = filtersc;
  const float seq = (float)((double)(usc - filtersc) / PA_LN2);
  if (!(seq >= hm.thrF1)) return;  // P > F1
  Pass2 o;
  o.t = r.t; o.h = r.h; o.flags = (seq >= hm.thrF2m) ? 0 : 1; o.filtersc = filtersc;
  unsigned long long idx = atomicAdd(a.out_count, 1ull);
  if (idx < a.out_cap) a.out[idx] = o;
}

// ---------------------------------------------------------------------------------------------
// Viterbi filter
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_max_s32(int v) { return __reduce_max_sync(PA_FULL, v); }

__device__ __forceinline__ float vit_score_from_xC(const DevHmm &hm, int xC, int xw_move) {
  if (xC >= 32767) return __int_as_float(0x7f800000);
  if (xC <= PA_NEG16) return __int_as_float(0xff800000);
  float sc = (float)xC + (float)xw_move - (float)hm.base_w;
  sc /= hm.scale_w;
  sc -= 3.0f;
  return sc;
}

// One (profile, target) pair per warp; DP state in registers. Lane l owns cells k = l*BT+1..(l+1)*BT
// (BT = the profile's class-rounded block size; padded cells carry -32768 scores). Tables [j][lane]:
// trA int4 {BM,MM,IM,DM}(k-1) | trI int2 {MI,II}(k) | trD int2 {MD,DD}(k-1) | ps int | em[x] int.
// Saturating signed-16 semantics are evaluated exactly in s32: VIADDMNMX fuses add + max with the
// -32768 floor; the upper saturation can only be reached in the row maximum, which ends the pass.
template <int BT>
__device__ __forceinline__ int vit_pair_reg(const DevHmm &hm, const int *tab, const uint8_t *p, int L, int xw_move, int lane) {
  constexpr int W = BT * 32;
  const int4 *trA = reinterpret_cast<const int4 *>(tab + hm.vit_off);
  const int2 *trI = reinterpret_cast<const int2 *>(trA + W);
  const int2 *trD = trI + W;
  const int *ps = reinterpret_cast<const int *>(trD + W);
  const int *em = ps + W;
  int Mx[BT], Ix[BT], Dx[BT];
#pragma unroll
  for (int j = 0; j < BT; j++) { Mx[j] = PA_NEG16; Ix[j] = PA_NEG16; Dx[j] = PA_NEG16; }
  const int xN = hm.base_w, xw_E = hm.xw_E;
  int xB = max(xN + xw_move, PA_NEG16), xJ = PA_NEG16, xC = PA_NEG16;
  const int pslast = ps[(BT - 1) * 32 + lane];
  for (int i = 0; i < L; i++) {
    const int *emx = em + (size_t)p[i] * W;
    int mL = __shfl_up_sync(PA_FULL, Mx[BT - 1], 1), iL = __shfl_up_sync(PA_FULL, Ix[BT - 1], 1),
        dL = __shfl_up_sync(PA_FULL, Dx[BT - 1], 1);
    if (lane == 0) mL = iL = dL = PA_NEG16;
    int rowmax = PA_NEG16;
#pragma unroll
    for (int j = BT - 1; j >= 0; j--) {
      const int c = j * 32 + lane;
      const int4 A = __ldg(trA + c);
      const int2 Iq = __ldg(trI + c);
      const int e = __ldg(emx + c);
      const int pm = j ? Mx[j > 0 ? j - 1 : 0] : mL, pi = j ? Ix[j > 0 ? j - 1 : 0] : iL, pd = j ? Dx[j > 0 ? j - 1 : 0] : dL;
      int sv = __viaddmax_s32(xB, A.x, PA_NEG16);
      sv = __viaddmax_s32(pm, A.y, sv);
      sv = __viaddmax_s32(pi, A.z, sv);
      sv = __viaddmax_s32(pd, A.w, sv);
      sv = __viaddmax_s32(sv, e, PA_NEG16);
      int iv = __viaddmax_s32(Mx[j], Iq.x, PA_NEG16);
      iv = __viaddmax_s32(Ix[j], Iq.y, iv);
      Mx[j] = sv;
      Ix[j] = iv;
      rowmax = max(rowmax, sv);
      if ((j & 3) == 0) asm volatile("" ::: "memory");  // bound load hoisting (register pressure)
    }
    const int xE = __reduce_max_sync(PA_FULL, rowmax);
    if (xE >= 32767) return 32767;
    int mcL = __shfl_up_sync(PA_FULL, Mx[BT - 1], 1);
    if (lane == 0) mcL = PA_NEGBIG;
    int aloc = PA_NEGBIG;
#pragma unroll
    for (int j = 0; j < BT; j++) {
      const int2 Dq = __ldg(trD + j * 32 + lane);
      const int mleft = j ? Mx[j > 0 ? j - 1 : 0] : mcL;
      aloc = max(max(mleft + Dq.x, aloc + Dq.y), PA_NEGBIG);
      Dx[j] = aloc;
    }
    int LA = aloc, LS = pslast;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a2 = __shfl_up_sync(PA_FULL, LA, o), s2 = __shfl_up_sync(PA_FULL, LS, o);
      if (lane >= o) {
        LA = max(LA, max(a2 + LS, PA_NEGBIG));
        LS = max(LS + s2, PA_NEGBIG);
      }
    }
    int din = __shfl_up_sync(PA_FULL, LA, 1);
    if (lane == 0) din = PA_NEGBIG;
#pragma unroll
    for (int j = 0; j < BT; j++) Dx[j] = __vimax3_s32(Dx[j], din + __ldg(ps + j * 32 + lane), PA_NEG16);
    xC = max(xC, max(xE + xw_E, PA_NEG16));
    xJ = max(xJ, max(xE + xw_E, PA_NEG16));
    xB = max(max(xJ + xw_move, xN + xw_move), PA_NEG16);
  }
  return xC;
}

struct VitArgs {
  const DevHmm *hmms;
  const int *tab;          // vit pool
  Targets tg;
  LenTabs lt;
  const Pass2 *in;         // sorted by profile rank; this launch handles [begin, end)
  unsigned long long begin, end;
  Pass2 *out;
  unsigned long long *out_count;
  unsigned long long out_cap;
  int all_pairs;           // score_all: run Viterbi on every input, write raw xC at the input index
  int *raw_out;
  const uint8_t *need;     // from vit16_kernel: 0 = the upper bound already failed the filter, skip (nullptr: run all)
};

// register budget per class: keep several warps per scheduler for the common (M <= 512) profiles
template <int BT>
static __global__ void __launch_bounds__(128, (BT <= 8 ? 5 : BT <= 16 ? 4 : BT <= 24 ? 3 : BT <= 32 ? 2 : 1)) vit_kernel(VitArgs a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (unsigned long long i = a.begin + (unsigned long long)blockIdx.x * wpb + wib; i < a.end; i += (unsigned long long)gridDim.x * wpb) {
    if (a.need && !a.need[i]) continue;
    const Pass2 r = a.in[i];
    const DevHmm &hm = a.hmms[r.h];
    if (!(r.flags & 1) && !a.all_pairs) {  // P <= F2 already: Viterbi not needed
      if (lane == 0) {
        unsigned long long idx = atomicAdd(a.out_count, 1ull);
        if (idx < a.out_cap) a.out[idx] = r;
      }
      continue;
    }
    const int L = (int)a.tg.len[r.t];
    const int xw_move = a.lt.xw_move[a.tg.lidx[r.t]];
    const int xC = vit_pair_reg<BT>(hm, a.tab, a.tg.dsq + a.tg.off[r.t], L, xw_move, lane);
    if (a.raw_out && lane == 0) a.raw_out[i] = xC;
    if (a.all_pairs) continue;
    const float vfsc = vit_score_from_xC(hm, xC, xw_move);
    const float seq = (float)((double)(vfsc - r.filtersc) / PA_LN2);
    if (seq >= hm.thrF2v && lane == 0) {
      unsigned long long idx = atomicAdd(a.out_count, 1ull);
      if (idx < a.out_cap) a.out[idx] = r;
    }
  }
}

// ---- counting sort of a Pass2 list by profile rank (groups pairs of one profile together) ----
static __global__ void hist_kernel(const Pass2 *in, unsigned long long n, const DevHmm *hmms, unsigned int *cnt) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[hmms[in[i].h].rank], 1u);
}
// single block: exclusive scan of cnt[0..nh) into start[] (also copied to cursor[]); start[nh] = total
static __global__ void scan_kernel(const unsigned int *cnt, int nh, unsigned int *start, unsigned int *cursor) {
  __shared__ unsigned int carry;
  __shared__ unsigned int buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nh; base += 1024) {
    const int i = base + threadIdx.x;
    unsigned int v = i < nh ? cnt[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nh) { start[i] = carry + buf[threadIdx.x] - v; cursor[i] = start[i]; }
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) start[nh] = carry;
}
static __global__ void scatter_kernel(const Pass2 *in, unsigned long long n, const DevHmm *hmms, unsigned int *cursor, Pass2 *out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n) {
    const Pass2 r = in[i];
    out[atomicAdd(&cursor[hmms[r.h].rank], 1u)] = r;
  }
}

// ---------------------------------------------------------------------------------------------
// Forward / Backward rows in scaled probability space (canonical order)
// ---------------------------------------------------------------------------------------------
struct XF { float tNN, tNB, tJJ, tJB, tCC, tCT, tEJ, tEC; };
struct Specials { float xE, xN, xJ, xB, xC; };

__device__ __forceinline__ XF xf_config(int Lcfg, int multihit) {
  XF x;
  float nj = multihit ? 1.0f : 0.0f;
  float pmove = __fdiv_rn(2.0f + nj, __fadd_rn(__fadd_rn((float)Lcfg, 2.0f), nj));
  float ploop = __fsub_rn(1.0f, pmove);
  x.tNN = x.tJJ = x.tCC = ploop;
  x.tNB = x.tJB = x.tCT = pmove;
  x.tEJ = multihit ? 0.5f : 0.0f;
  x.tEC = multihit ? 0.5f : 1.0f;
  return x;
}

__device__ __forceinline__ float lane_sum(float s) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(PA_FULL, s, o));
  return s;
}

struct FwdTabs {
  const float4 *trA;  // {BM[k-1], MM[k-1], IM[k-1], DM[k-1]}
  const float4 *trB;  // {MD[k-1], DD[k-1], MI[k], II[k]}
  const float4 *bwA;  // {MM[k], MI[k], MD[k], IM[k]}
  const float4 *bwB;  // {II[k], DM[k], DD[k], BM[k-1]}
  const float *em;    // [x][W]
  int B, W, M;
};
__device__ __forceinline__ FwdTabs fwd_tabs(const DevHmm &hm, const float *pool) {
  FwdTabs t;
  t.B = hm.B; t.W = hm.B * 32; t.M = hm.M;
  t.trA = reinterpret_cast<const float4 *>(pool + hm.fwd_off);
  t.trB = t.trA + t.W;
  t.bwA = t.trB + t.W;
  t.bwB = t.bwA + t.W;
  t.em = reinterpret_cast<const float *>(t.bwB + t.W);
  return t;
}

// One forward row. Arrays are [j*32+lane]; cells beyond M stay 0. wP: scratch row. Returns xE.
static __device__ float fwd_row(const FwdTabs &T, int x, const float *Mp, const float *Ip, const float *Dp, float xBp,
                         float *Mc, float *Ic, float *Dc, float *wP, int lane) {
  const int B = T.B;
  const float *em = T.em + (size_t)x * T.W;
  const int nv = min(B, max(0, T.M - lane * B));  // valid cells of this lane
  const int ilast = (B - 1) * 32 + lane;
  float mL = __shfl_up_sync(PA_FULL, Mp[ilast], 1), iL = __shfl_up_sync(PA_FULL, Ip[ilast], 1),
        dL = __shfl_up_sync(PA_FULL, Dp[ilast], 1);
  if (lane == 0) mL = iL = dL = 0.0f;
  for (int j = 0; j < nv; j++) {
    const int c = j * 32 + lane;
    const float4 A = T.trA[c], Bq = T.trB[c];
    const float pm = j ? Mp[c - 32] : mL, pi = j ? Ip[c - 32] : iL, pd = j ? Dp[c - 32] : dL;
    float t = __fmul_rn(xBp, A.x);
    t = __fadd_rn(t, __fmul_rn(pm, A.y));
    t = __fadd_rn(t, __fmul_rn(pi, A.z));
    t = __fadd_rn(t, __fmul_rn(pd, A.w));
    Mc[c] = __fmul_rn(t, em[c]);
    Ic[c] = __fadd_rn(__fmul_rn(Mp[c], Bq.z), __fmul_rn(Ip[c], Bq.w));
  }
  __syncwarp();
  float mcL = __shfl_up_sync(PA_FULL, Mc[ilast], 1);
  if (lane == 0) mcL = 0.0f;
  float LA = 0.0f, LP = 1.0f;
  for (int j = 0; j < nv; j++) {
    const int c = j * 32 + lane;
    const float4 Bq = T.trB[c];
    const float a = __fmul_rn(j ? Mc[c - 32] : mcL, Bq.x);
    if (j == 0) { LA = a; LP = Bq.y; }
    else { LA = __fadd_rn(a, __fmul_rn(LA, Bq.y)); LP = __fmul_rn(LP, Bq.y); }
    Dc[c] = LA;
    wP[c] = LP;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float a2 = __shfl_up_sync(PA_FULL, LA, o), p2 = __shfl_up_sync(PA_FULL, LP, o);
    if (lane >= o) { LA = __fadd_rn(LA, __fmul_rn(LP, a2)); LP = __fmul_rn(LP, p2); }
  }
  float cin = __shfl_up_sync(PA_FULL, LA, 1);
  if (lane == 0) cin = 0.0f;
  float s = 0.0f;
  for (int j = 0; j < nv; j++) {
    const int c = j * 32 + lane;
    const float d = __fadd_rn(Dc[c], __fmul_rn(wP[c], cin));
    Dc[c] = d;
    s = __fadd_rn(s, __fadd_rn(Mc[c], d));
  }
  __syncwarp();
  return lane_sum(s);
}

__device__ __forceinline__ void scale_row(float *Mc, float *Ic, float *Dc, int B, int lane, int e) {
  for (int j = 0; j < B; j++) {
    const int c = j * 32 + lane;
    Mc[c] = scalbnf(Mc[c], -e); Ic[c] = scalbnf(Ic[c], -e); Dc[c] = scalbnf(Dc[c], -e);
  }
}

// Forward parser over p[0..L-1]; sm: 7 rows of W floats (zero-initialised here). Optionally stores
// per-row specials/exponents (sp/sexp index 0..L). Returns exponent, *mant = xC*tCT.
static __device__ int fwd_parser(const FwdTabs &T, const uint8_t *p, int L, int Lcfg, int multihit, float *sm, int lane,
                          float *mant, Specials *sp, int *sexp) {
  const int W = T.W;
  float *Mp = sm, *Ip = sm + W, *Dp = sm + 2 * W, *Mc = sm + 3 * W, *Ic = sm + 4 * W, *Dc = sm + 5 * W, *wP = sm + 6 * W;
  for (int c = lane; c < 7 * W; c += 32) sm[c] = 0.0f;
  __syncwarp();
  const XF xf = xf_config(Lcfg, multihit);
  float xN = 1.0f, xB = xf.tNB, xE = 0.0f, xJ = 0.0f, xC = 0.0f;
  int etot = 0;
  if (sp && lane == 0) { Specials s0; s0.xE = 0; s0.xN = xN; s0.xJ = 0; s0.xB = xB; s0.xC = 0; sp[0] = s0; sexp[0] = 0; }
  for (int i = 0; i < L; i++) {
    xE = fwd_row(T, p[i], Mp, Ip, Dp, xB, Mc, Ic, Dc, wP, lane);
    xN = __fmul_rn(xN, xf.tNN);
    xJ = __fadd_rn(__fmul_rn(xJ, xf.tJJ), __fmul_rn(xE, xf.tEJ));
    xC = __fadd_rn(__fmul_rn(xC, xf.tCC), __fmul_rn(xE, xf.tEC));
    xB = __fadd_rn(__fmul_rn(xJ, xf.tJB), __fmul_rn(xN, xf.tNB));
    if (xE > PA_RESCALE) {
      const int e = ilogbf(xE);
      scale_row(Mc, Ic, Dc, T.B, lane, e);
      xN = scalbnf(xN, -e); xJ = scalbnf(xJ, -e); xC = scalbnf(xC, -e); xB = scalbnf(xB, -e); xE = scalbnf(xE, -e);
      etot += e;
      __syncwarp();
    }
    if (sp && lane == 0) { Specials s; s.xE = xE; s.xN = xN; s.xJ = xJ; s.xB = xB; s.xC = xC; sp[i + 1] = s; sexp[i + 1] = etot; }
    float *t;
    t = Mp; Mp = Mc; Mc = t;
    t = Ip; Ip = Ic; Ic = t;
    t = Dp; Dp = Dc; Dc = t;
  }
  *mant = __fmul_rn(xC, xf.tCT);
  return etot;
}

struct FwdArgs {
  const DevHmm *hmms;
  const float *tab;
  Targets tg;
  const Pass2 *in;
  unsigned long long n_in;
  Pass4 *out;
  unsigned long long *out_count;
  unsigned long long out_cap;
  int maxB;
  int all_pairs;       // score_all: emit every input with its forward result, in input order
};

// ---------------------------------------------------------------------------------------------
// Forward filter, block size known at compile time (B = ceil(M/32), the canonical lane block).
// Same arithmetic in the same order as fwd_row/fwd_parser above (and oracle/p7_filters.c), so the
// scores are bit-identical; what changes is where the row lives and how much the compiler can overlap:
//   B <= 32  M/I/D/P of the lane's block in registers (4B floats), loops fully unrolled
//   B  > 32  same code over shared-memory rows [j][lane] (conflict-free), 4 rows per warp instead of 7
// Pairs arrive sorted by profile, one launch per distinct B.
template <int B>
struct FwdRegRows {
  float m[B], i[B], d[B], p[B];
  __device__ __forceinline__ FwdRegRows(float *, int) {}
  __device__ __forceinline__ float &M(int j) { return m[j]; }
  __device__ __forceinline__ float &I(int j) { return i[j]; }
  __device__ __forceinline__ float &D(int j) { return d[j]; }
  __device__ __forceinline__ float &P(int j) { return p[j]; }
};
template <int B>
struct FwdSmemRows {
  float *b;
  __device__ __forceinline__ FwdSmemRows(float *base, int lane) : b(base + lane) {}
  __device__ __forceinline__ float &M(int j) { return b[j * 32]; }
  __device__ __forceinline__ float &I(int j) { return b[(B + j) * 32]; }
  __device__ __forceinline__ float &D(int j) { return b[(2 * B + j) * 32]; }
  __device__ __forceinline__ float &P(int j) { return b[(3 * B + j) * 32]; }
};

template <int B, class Rows>
__device__ __forceinline__ int fwd_pair_t(const FwdTabs &T, const uint8_t *p, int L, int lane, float *smem_rows, float *mant) {
  constexpr int W = B * 32;
  Rows R(smem_rows, lane);
#pragma unroll
  for (int j = 0; j < B; j++) { R.M(j) = 0.0f; R.I(j) = 0.0f; R.D(j) = 0.0f; R.P(j) = 0.0f; }
  const int nv = min(B, max(0, T.M - lane * B));  // valid cells of this lane
  const XF xf = xf_config(L, 1);
  float xN = 1.0f, xB = xf.tNB, xE = 0.0f, xJ = 0.0f, xC = 0.0f;
  int etot = 0;
  for (int i = 0; i < L; i++) {
    const float *em = T.em + (size_t)p[i] * W;
    float mL = __shfl_up_sync(PA_FULL, R.M(B - 1), 1), iL = __shfl_up_sync(PA_FULL, R.I(B - 1), 1),
          dL = __shfl_up_sync(PA_FULL, R.D(B - 1), 1);
    if (lane == 0) mL = iL = dL = 0.0f;
    // M and I, in place: descending j reads the previous row's j-1 before it is overwritten
#pragma unroll
    for (int j = B - 1; j >= 0; j--) {
      if (j < nv) {
        const int c = j * 32 + lane;
        const float4 A = __ldg(T.trA + c);
        const float4 Bq = __ldg(T.trB + c);
        const float e = __ldg(em + c);
        const float pm = j ? R.M(j > 0 ? j - 1 : 0) : mL, pi = j ? R.I(j > 0 ? j - 1 : 0) : iL, pd = j ? R.D(j > 0 ? j - 1 : 0) : dL;
        float t = __fmul_rn(xB, A.x);
        t = __fadd_rn(t, __fmul_rn(pm, A.y));
        t = __fadd_rn(t, __fmul_rn(pi, A.z));
        t = __fadd_rn(t, __fmul_rn(pd, A.w));
        const float ni = __fadd_rn(__fmul_rn(R.M(j), Bq.z), __fmul_rn(R.I(j), Bq.w));
        R.M(j) = __fmul_rn(t, e);
        R.I(j) = ni;
      }
    }
    // D chain: in-lane serial prefix, scan across lanes, fix-up
    float mcL = __shfl_up_sync(PA_FULL, R.M(B - 1), 1);
    if (lane == 0) mcL = 0.0f;
    float LA = 0.0f, LP = 1.0f;
#pragma unroll
    for (int j = 0; j < B; j++) {
      if (j < nv) {
        const float4 Bq = __ldg(T.trB + j * 32 + lane);
        const float a = __fmul_rn(j ? R.M(j > 0 ? j - 1 : 0) : mcL, Bq.x);
        if (j == 0) { LA = a; LP = Bq.y; }
        else { LA = __fadd_rn(a, __fmul_rn(LA, Bq.y)); LP = __fmul_rn(LP, Bq.y); }
        R.D(j) = LA;
        R.P(j) = LP;
      }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float a2 = __shfl_up_sync(PA_FULL, LA, o), p2 = __shfl_up_sync(PA_FULL, LP, o);
      if (lane >= o) { LA = __fadd_rn(LA, __fmul_rn(LP, a2)); LP = __fmul_rn(LP, p2); }
    }
    float cin = __shfl_up_sync(PA_FULL, LA, 1);
    if (lane == 0) cin = 0.0f;
    float sacc = 0.0f;
#pragma unroll
    for (int j = 0; j < B; j++) {
      if (j < nv) {
        const float d = __fadd_rn(R.D(j), __fmul_rn(R.P(j), cin));
        R.D(j) = d;
        sacc = __fadd_rn(sacc, __fadd_rn(R.M(j), d));
      }
    }
    xE = lane_sum(sacc);
    xN = __fmul_rn(xN, xf.tNN);
    xJ = __fadd_rn(__fmul_rn(xJ, xf.tJJ), __fmul_rn(xE, xf.tEJ));
    xC = __fadd_rn(__fmul_rn(xC, xf.tCC), __fmul_rn(xE, xf.tEC));
    xB = __fadd_rn(__fmul_rn(xJ, xf.tJB), __fmul_rn(xN, xf.tNB));
    if (xE > PA_RESCALE) {
      const int e = ilogbf(xE);
#pragma unroll
      for (int j = 0; j < B; j++) { R.M(j) = scalbnf(R.M(j), -e); R.I(j) = scalbnf(R.I(j), -e); R.D(j) = scalbnf(R.D(j), -e); }
      xN = scalbnf(xN, -e); xJ = scalbnf(xJ, -e); xC = scalbnf(xC, -e); xB = scalbnf(xB, -e); xE = scalbnf(xE, -e);
      etot += e;
    }
  }
  *mant = __fmul_rn(xC, xf.tCT);
  return etot;
}

struct FwdTArgs {
  FwdArgs f;
  unsigned long long begin, end;  // range of f.in handled by this launch (all pairs share the block size B)
};

template <int B>
static __global__ void __launch_bounds__(128) fwd_kernel_t(FwdTArgs a) {
  extern __shared__ __align__(16) float s_fwdt[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float *rows = s_fwdt + (size_t)wib * 4 * B * 32;  // used only when B > 32
  for (unsigned long long i = a.begin + (unsigned long long)blockIdx.x * wpb + wib; i < a.end; i += (unsigned long long)gridDim.x * wpb) {
    const Pass2 r = a.f.in[i];
    const DevHmm &hm = a.f.hmms[r.h];
    const FwdTabs T = fwd_tabs(hm, a.f.tab);
    const int L = (int)a.f.tg.len[r.t];
    float mant;
    int e;
    if constexpr (B <= 32) e = fwd_pair_t<B, FwdRegRows<B>>(T, a.f.tg.dsq + a.f.tg.off[r.t], L, lane, rows, &mant);
    else e = fwd_pair_t<B, FwdSmemRows<B>>(T, a.f.tg.dsq + a.f.tg.off[r.t], L, lane, rows, &mant);
    const float fwdsc = (float)((double)e * PA_LN2 + log((double)mant));
    const float seq = (float)((double)(fwdsc - r.filtersc) / PA_LN2);
    if (lane == 0 && (a.f.all_pairs || seq >= hm.thrF3)) {
      Pass4 o;
      o.t = r.t; o.h = r.h; o.pad = 0; o.filtersc = r.filtersc; o.fexp = e; o.fmant = mant;
      if (a.f.all_pairs) a.f.out[i] = o;
      else {
        unsigned long long idx = atomicAdd(a.f.out_count, 1ull);
        if (idx < a.f.out_cap) a.f.out[idx] = o;
      }
    }
  }
}

}  // namespace pa

// ---------------------------------------------------------------------------------------------
// integer-pipe microbenchmark: independent VIADDMNMX.S16x2(.RELU) / VIMNMX3.S16x2 chains.
// Gives the measured 16-bit lane-op peak the filter rooflines are quoted against.
// ---------------------------------------------------------------------------------------------
namespace pa {
template <int MODE>
static __global__ void __launch_bounds__(256) dpx_microbench_kernel(uint32_t *out, int iters, uint32_t seed) {
  uint32_t v[8], u[8], d = seed ^ threadIdx.x, d2 = (seed >> 3) ^ threadIdx.x, m = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = (threadIdx.x * 2654435761u + i) & 0x00ff00ffu; u[i] = v[i] ^ 0x00110011u; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) v[i] = __viaddmax_s16x2_relu(v[i], d, 0u);          // SSV cell op
      if (MODE == 1) v[i] = __vimax3_s16x2(v[i], d, m);                  // 3-input max
      if (MODE == 2) { v[i] = __viaddmax_s16x2_relu(v[i], d, 0u); if (i & 1) m = __vimax3_s16x2(m, v[i], v[i - 1]); }  // SSV mix
      if (MODE == 3) v[i] = (uint32_t)__viaddmax_s32((int)v[i], (int)d, -32768);  // Viterbi op
      if (MODE == 4 || MODE == 5) {  // fp16x2 SSV cell op of msv_tmem_kernel: HFMA2.RELU (FMA pipe)
        asm("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(0x3c003c00u), "r"(d));
        if (MODE == 5 && (i & 1)) m = __vimax3_s16x2(m, v[i], v[i - 1]);  // + running max (ALU pipe)
      }
    }
    if (MODE == 6 || MODE == 7) {
      // can the ALU pipe's DPX formulation run NEXT TO the FMA pipe's fp16 formulation?  6: even warps HFMA2.RELU chains, odd
      // warps the DPX SSV mix (8 VIADDMNMX.S16x2.RELU + 4 VIMNMX3.S16x2); 7: both in every warp, on independent registers
      const bool fma_warp = MODE == 7 || ((threadIdx.x >> 5) & 1) == 0, dpx_warp = MODE == 7 || !fma_warp;
      if (fma_warp) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(0x3c003c00u), "r"(d));
      }
      if (dpx_warp) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          u[i] = __viaddmax_s16x2_relu(u[i], d2, 0u);
          if (i & 1) m = __vimax3_s16x2(m, u[i], u[i - 1]);
        }
      }
      d2 += 0x00010001u;
    }
    if (MODE >= 4) d ^= 0x80008000u; else d += 0x00010001u;
  }
  if (MODE >= 6) {
#pragma unroll
    for (int i = 0; i < 8; i++) m ^= u[i];
  }
  uint32_t r = m;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= v[i];
  if (r == 0x12345678u) out[0] = r;
}
}  // namespace pa
