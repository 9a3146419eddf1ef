// probe_sm100.cu -- standalone microbenchmarks behind the msv_kernel design (run on the B200 box):
//   T1  HFMA2.RELU issue rate (fp16x2 SSV cell update on the FMA pipe)
//   T2  SSV mix: 8 HFMA2.RELU + 4 VIMNMX3.S16x2
//   T3  tcgen05.ld 32x32b.x8 throughput (TMEM -> registers) for 4/8/16 warps per SM
//   T4  emulated SSV row loop, emission rows from TMEM   (cells/clk/SM)
//   T5  emulated SSV row loop, emission rows from shared memory (LDS.128)
//   T6  rows alternate between TMEM and shared memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_sm100 probe_sm100.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t hfma2_relu(uint32_t a, uint32_t one, uint32_t c) {
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) alu_kernel(uint32_t *out, int iters, uint32_t seed) {
  uint32_t v[8], d = seed ^ (threadIdx.x & 3), m = 0;
  const uint32_t one = 0x3c003c00u;
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) v[i] = hfma2_relu(v[i], one, d);
      if (MODE == 1) { v[i] = hfma2_relu(v[i], one, d); if (i & 1) m = __vimax3_s16x2(m, v[i], v[i - 1]); }
      if (MODE == 2) { v[i] = __viaddmax_s16x2_relu(v[i], d, 0u); if (i & 1) m = __vimax3_s16x2(m, v[i], v[i - 1]); }
    }
    d ^= 0x80008000u;
  }
  uint32_t r = m;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= v[i];
  if (r == 0x12345678u) out[0] = r;
}

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot, int ncols_pow2) {
  // warp 0 allocates; everyone reads the base address
  if (threadIdx.x < 32) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(ncols_pow2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, int ncols) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}

// T3: pure tcgen05.ld throughput
__global__ void __launch_bounds__(512) tmem_bw_kernel(uint32_t *out, int iters) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc_all(&slot, 512);
  const int warp = threadIdx.x >> 5;
  const uint32_t lanebase = base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t init[8];
#pragma unroll
  for (int i = 0; i < 8; i++) init[i] = threadIdx.x * 8 + i;
  if (warp < 4) {
    for (int c = 0; c < 512; c += 8) tmem_st8(lanebase + c, init);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t acc = 0;
  uint32_t col = (warp * 40) & 511;
  for (int it = 0; it < iters; it++) {
    uint32_t a[8], b[8];
    tmem_ld8(lanebase + (col & 504), a);
    tmem_ld8(lanebase + ((col + 64) & 504), b);
    tmem_wait_ld8(a);
    tmem_wait_ld8(b);
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] + b[i];
    col += 8;
  }
  if (acc == 0x12345678u) out[0] = acc;
  tmem_free(base, 512);
}

// T4/T5/T6: emulated SSV inner loop, Q = 8 words per lane (M <= 511), 24 table rows
template <int SRC, int FP16>
__global__ void __launch_bounds__(512) ssv_emul_kernel(uint32_t *out, int nrows_total, uint32_t seed) {
  constexpr int Q = 8, NR = 24;
  __shared__ uint32_t slot;
  extern __shared__ __align__(16) uint32_t s_tab[];  // NR * Q*32 words
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = 0, lanebase = 0;
  // table: small negative deltas so the recurrence stays bounded
  for (int i = threadIdx.x; i < NR * Q * 32; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ seed;
    int d0 = (int)(h & 7) - 5, d1 = (int)((h >> 8) & 7) - 5;
    if (FP16) {
      __half2 v = __halves2half2(__int2half_rn(d0), __int2half_rn(d1));
      s_tab[i] = *reinterpret_cast<uint32_t *>(&v);
    } else
      s_tab[i] = (uint32_t)(uint16_t)(int16_t)d0 | ((uint32_t)(uint16_t)(int16_t)d1 << 16);
  }
  __syncthreads();
  if (SRC != 1) {
    base = tmem_alloc_all(&slot, 256);
    lanebase = base + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
      for (int x = 0; x < NR; x++) {
        uint32_t r[8];
        const uint4 v0 = *reinterpret_cast<const uint4 *>(s_tab + x * Q * 32 + lane * 4);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(s_tab + x * Q * 32 + 128 + lane * 4);
        r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
        tmem_st8(lanebase + x * 8, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t one = 0x3c003c00u, zero = seed >> 31;
  const int lanem1 = (lane + 31) & 31;
  uint32_t w[Q], m = 0;
#pragma unroll
  for (int q = 0; q < Q; q++) w[q] = 0;
  uint32_t rs = seed * (warp + 1) + blockIdx.x;
  auto next_row = [&]() { rs = rs * 1664525u + 1013904223u; return (rs >> 16) % NR; };
  auto load = [&](int x, int which, uint32_t (&d)[8]) {
    const bool from_tmem = SRC == 0 || (SRC == 2 && (which & 1));
    if (from_tmem) tmem_ld8(lanebase + x * 8, d);
    else {
      const uint4 v0 = *reinterpret_cast<const uint4 *>(s_tab + x * Q * 32 + lane * 4);
      const uint4 v1 = *reinterpret_cast<const uint4 *>(s_tab + x * Q * 32 + 128 + lane * 4);
      d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w; d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
    }
  };
  auto wait = [&](int which, uint32_t (&d)[8]) {
    const bool from_tmem = SRC == 0 || (SRC == 2 && (which & 1));
    if (from_tmem) tmem_wait_ld8(d);
  };
  auto row = [&](const uint32_t (&d)[8]) {
    const uint32_t tt = __shfl_sync(FULL, w[Q - 1], lanem1);
    const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
    if (FP16) {
#pragma unroll
      for (int q = Q - 1; q >= 1; q--) w[q] = hfma2_relu(w[q - 1], one, d[q]);
      w[0] = hfma2_relu(carry, one, d[0]);
    } else {
#pragma unroll
      for (int q = Q - 1; q >= 1; q--) w[q] = __viaddmax_s16x2(w[q - 1], d[q], zero);
      w[0] = __viaddmax_s16x2(carry, d[0], zero);
    }
#pragma unroll
    for (int q = 0; q + 1 < Q; q += 2) m = __vimax3_s16x2(m, w[q], w[q + 1]);
  };
  uint32_t da[8], db[8];
  load(next_row(), 0, da);
  for (int r = 0; r < nrows_total; r += 2) {
    wait(0, da);
    load(next_row(), 1, db);
    row(da);
    wait(1, db);
    load(next_row(), 0, da);
    row(db);
  }
  wait(0, da);
  uint32_t acc = m ^ da[0];
#pragma unroll
  for (int q = 0; q < Q; q++) acc ^= w[q];
  if (acc == 0x12345678u) out[0] = acc;
  if (SRC != 1) tmem_free(base, 256);
}


// T7: SSV emulation v2 -- residues from global memory, row offsets broadcast by shuffle, rows fetched in
// groups of G (double-buffered): loads of group g+1 are in flight while group g is computed.
template <int G, int FP16, int MINB>
__global__ void __launch_bounds__(512, MINB) ssv_emul2_kernel(uint32_t *out, const uint8_t *seq, int L, int ntargets_per_warp, uint32_t seed) {
  constexpr int Q = 8, NR = 24;
  __shared__ uint32_t slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = tmem_alloc_all(&slot, 256);
  const uint32_t lanebase = base + ((uint32_t)((warp & 3) * 32) << 16);
  if (warp < 4) {
    for (int x = 0; x < NR; x++) {
      uint32_t r[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint32_t h = ((x * 256 + q * 32 + lane) * 2654435761u) ^ seed;
        int d0 = (int)(h & 7) - 5, d1 = (int)((h >> 8) & 7) - 5;
        if (FP16) {
          __half2 v = __halves2half2(__int2half_rn(d0), __int2half_rn(d1));
          r[q] = *reinterpret_cast<uint32_t *>(&v);
        } else
          r[q] = (uint32_t)(uint16_t)(int16_t)d0 | ((uint32_t)(uint16_t)(int16_t)d1 << 16);
      }
      tmem_st8(lanebase + x * 8, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t one = 0x3c003c00u, zero = seed >> 31;
  const int lanem1 = (lane + 31) & 31;
  uint32_t acc = 0;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  for (int t = 0; t < ntargets_per_warp; t++) {
    const uint8_t *p = seq + ((size_t)(gw * 7 + t) % 4096) * 64;
    uint32_t w[Q], m0 = 0, m1 = 0;
#pragma unroll
    for (int q = 0; q < Q; q++) w[q] = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
      const uint32_t curoff = (i0 + lane < L) ? (uint32_t)p[i0 + lane] * 8u : 0u;
      const int nr = min(32, L - i0);  // multiple of G in this probe
      uint32_t buf[2][G][8];
#pragma unroll
      for (int g = 0; g < G; g++) tmem_ld8(lanebase + __shfl_sync(FULL, curoff, g), buf[0][g]);
#pragma unroll 1
      for (int r = 0; r < nr; r += 2 * G) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
          const int rr = r + half * G;
          if (rr < nr) {
            if (rr + G < nr) {
#pragma unroll
              for (int g = 0; g < G; g++) tmem_ld8(lanebase + __shfl_sync(FULL, curoff, rr + G + g), buf[half ^ 1][g]);
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
              uint32_t(&d)[8] = buf[half][g];
              tmem_wait_ld8(d);
              const uint32_t tt = __shfl_sync(FULL, w[Q - 1], lanem1);
              const uint32_t carry = __byte_perm(tt, w[Q - 1], 0x5432);
              if (FP16) {
#pragma unroll
                for (int q = Q - 1; q >= 1; q--) w[q] = hfma2_relu(w[q - 1], one, d[q]);
                w[0] = hfma2_relu(carry, one, d[0]);
              } else {
#pragma unroll
                for (int q = Q - 1; q >= 1; q--) w[q] = __viaddmax_s16x2(w[q - 1], d[q], zero);
                w[0] = __viaddmax_s16x2(carry, d[0], zero);
              }
              m0 = __vimax3_s16x2(m0, w[0], w[1]);
              m1 = __vimax3_s16x2(m1, w[2], w[3]);
              m0 = __vimax3_s16x2(m0, w[4], w[5]);
              m1 = __vimax3_s16x2(m1, w[6], w[7]);
            }
          }
        }
      }
    }
    acc += __reduce_max_sync(FULL, __vmaxs2(m0, m1));
  }
  if (acc == 0x12345678u) out[0] = acc;
  tmem_free(base, 256);
}

// T8: LDTM issue cost by width and address source.  W = columns per load (1,8,16,32); ADDR: 0 = loop-invariant
// uniform address, 1 = address from a per-iteration warp shuffle (vector reg -> R2UR), 2 = shuffle only (no LDTM)
template <int W, int ADDR>
__global__ void __launch_bounds__(512) ldtm_cost_kernel(uint32_t *out, int iters) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc_all(&slot, 512);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lanebase = base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t myaddr = lanebase + ((lane * 8) & 255);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t ad = lanebase + u * 32;
      if (ADDR >= 1) ad = __shfl_sync(FULL, myaddr, (it + u) & 31);
      if (ADDR == 2) { acc ^= ad; continue; }
      if (W == 1) {
        uint32_t r0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(ad) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0)::"memory");
        acc ^= r0;
      } else if (W == 8) {
        uint32_t a[8];
        tmem_ld8(ad, a);
        tmem_wait_ld8(a);
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= a[i];
      } else if (W == 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(ad)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
#pragma unroll
        for (int i = 0; i < 16; i++) acc ^= r[i];
      }
    }
    myaddr += 0;
  }
  if (acc == 0x12345678u) out[0] = acc;
  tmem_free(base, 512);
}

template <class F>
static float time_ms(F f, int reps = 4) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = clk_khz * 1e3;
  printf("%s: %d SMs, %.0f MHz\n", prop.name, sms, clk / 1e6);
  uint32_t *out;
  cudaMalloc(&out, 64);
  {
    const int iters = 4096, blocks = sms * 8;
    float t0 = time_ms([&] { alu_kernel<0><<<blocks, 256>>>(out, iters, 7u); });
    float t1 = time_ms([&] { alu_kernel<1><<<blocks, 256>>>(out, iters, 7u); });
    float t2 = time_ms([&] { alu_kernel<2><<<blocks, 256>>>(out, iters, 7u); });
    auto rate = [&](float ms, double per_iter) { return (double)blocks * 256 * iters * per_iter / (ms * 1e-3) / sms / clk; };
    printf("T1 HFMA2.RELU                     : %.1f thread-instr/clk/SM\n", rate(t0, 8));
    printf("T2 8 HFMA2.RELU + 4 VIMNMX3.S16x2  : %.1f thread-instr/clk/SM = %.1f cells/clk/SM\n", rate(t1, 12), rate(t1, 16));
    printf("T2' 8 VIADDMNMX.RELU + 4 VIMNMX3   : %.1f thread-instr/clk/SM = %.1f cells/clk/SM\n", rate(t2, 12), rate(t2, 16));
  }
  for (int threads : {128, 256, 512}) {
    const int iters = 20000;
    float t = time_ms([&] { tmem_bw_kernel<<<sms, threads>>>(out, iters); });
    double bytes = (double)sms * threads * iters * 64.0;
    printf("T3 tcgen05.ld 32x32b.x8, %2d warps/SM : %.1f B/clk/SM (%.2f TB/s chip)\n", threads / 32, bytes / (t * 1e-3) / sms / clk,
           bytes / (t * 1e-3) / 1e12);
  }
  const size_t smem = 24 * 8 * 32 * 4;
  for (int threads : {256, 512}) {
    const int nrows = 40000;
    auto run = [&](auto kern, const char *name) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float t = time_ms([&] { kern<<<sms, threads, smem>>>(out, nrows, 12345u); });
      double cells = (double)sms * (threads / 32) * nrows * 512.0;
      printf("%s, %2d warps/SM : %.1f cells/clk/SM (%.2f T cells/s)\n", name, threads / 32, cells / (t * 1e-3) / sms / clk, cells / (t * 1e-3) / 1e12);
    };
    run(ssv_emul_kernel<0, 1>, "T4 SSV emul  TMEM rows, fp16 HFMA2.RELU ");
    run(ssv_emul_kernel<0, 0>, "T4 SSV emul  TMEM rows, s16 VIADDMNMX   ");
    run(ssv_emul_kernel<1, 1>, "T5 SSV emul  smem rows, fp16 HFMA2.RELU ");
    run(ssv_emul_kernel<1, 0>, "T5 SSV emul  smem rows, s16 VIADDMNMX   ");
    run(ssv_emul_kernel<2, 1>, "T6 SSV emul  alternating, fp16          ");
    run(ssv_emul_kernel<2, 0>, "T6 SSV emul  alternating, s16           ");
  }

  {
    uint8_t *seq;
    cudaMalloc(&seq, 4096 * 64);
    uint8_t *h = (uint8_t *)malloc(4096 * 64);
    uint32_t s = 1;
    for (int i = 0; i < 4096 * 64; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 16) % 24; }
    cudaMemcpy(seq, h, 4096 * 64, cudaMemcpyHostToDevice);
    const int L = 48, nt = 2000;
    auto run2 = [&](auto kern, int blocks_per_sm, const char *name) {
      float t = time_ms([&] { kern<<<sms * blocks_per_sm, 512>>>(out, seq, L, nt, 12345u); });
      double cells = (double)sms * blocks_per_sm * 16 * nt * L * 512.0;
      printf("%s, %2d warps/SM : %.1f cells/clk/SM (%.2f T cells/s)\n", name, 16 * blocks_per_sm, cells / (t * 1e-3) / sms / clk, cells / (t * 1e-3) / 1e12);
    };
    run2(ssv_emul2_kernel<1, 1, 1>, 1, "T7 emul2 TMEM fp16 G=1");
    run2(ssv_emul2_kernel<2, 1, 1>, 1, "T7 emul2 TMEM fp16 G=2");
    run2(ssv_emul2_kernel<4, 1, 1>, 1, "T7 emul2 TMEM fp16 G=4");
    run2(ssv_emul2_kernel<1, 1, 2>, 2, "T7 emul2 TMEM fp16 G=1");
    run2(ssv_emul2_kernel<2, 1, 2>, 2, "T7 emul2 TMEM fp16 G=2");
    run2(ssv_emul2_kernel<4, 1, 2>, 2, "T7 emul2 TMEM fp16 G=4");
    run2(ssv_emul2_kernel<2, 0, 1>, 1, "T7 emul2 TMEM s16  G=2");
    run2(ssv_emul2_kernel<2, 0, 2>, 2, "T7 emul2 TMEM s16  G=2");
    cudaFree(seq);
    free(h);
  }

  {
    const int iters = 4000;
    auto run3 = [&](auto kern, int threads, const char *name, int W) {
      float t = time_ms([&] { kern<<<sms, threads>>>(out, iters); });
      double n = (double)(threads / 32) * iters * 4.0;   // warp-level loads per SM
      double clks = t * 1e-3 * clk;
      printf("%s, %2d warps/SM : %.2f clk/SM per warp-load (%.1f clk per SMSP), %.0f B/clk/SM\n", name, threads / 32, clks / n, 4 * clks / n, n * W * 128.0 / clks);
    };
    for (int th : {256, 512}) {
      run3(ldtm_cost_kernel<1, 0>, th, "T8 LDTM.x1  uniform addr", 1);
      run3(ldtm_cost_kernel<8, 0>, th, "T8 LDTM.x8  uniform addr", 8);
      run3(ldtm_cost_kernel<16, 0>, th, "T8 LDTM.x16 uniform addr", 16);
      run3(ldtm_cost_kernel<1, 1>, th, "T8 LDTM.x1  shfl+R2UR   ", 1);
      run3(ldtm_cost_kernel<8, 1>, th, "T8 LDTM.x8  shfl+R2UR   ", 8);
      run3(ldtm_cost_kernel<16, 1>, th, "T8 LDTM.x16 shfl+R2UR   ", 16);
      run3(ldtm_cost_kernel<8, 2>, th, "T8 SHFL only            ", 0);
    }
  }
  cudaFree(out);
  return 0;
}
