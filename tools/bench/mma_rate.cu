// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for different N and A-operand layouts, operands in SMEM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sketchedit_b200/csrc tools/bench/mma_rate.cu -o build/mma_rate
#include <cstdio>
#include <cstdlib>
#include "se_tc_device.cuh"
using namespace se;

// mode 0: A SW128 K-major (rows 128 B, SBO 1024), B SW128      mode 1: A no-swizzle (LBO, SBO given), B SW128
// mode 2: A SW64, B SW64
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ld_mode: 0 no epilogue traffic; 16 / 32: four (or eight) extra warps stream tcgen05.ld.x16 / .x32 of the OTHER accumulator half
__global__ void __launch_bounds__(384, 1) mma_rate(int N, int reps, int mode, int lbo, int sbo, int ld_mode, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand area so no NaN traps / denormal effects (64 KB A region + 64 KB B region)
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  for (int i = threadIdx.x; i < 32768; i += 384) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sA = base, sB = base + 65536;
    uint32_t a_hi, a_lo0, b_hi;
    if (mode == 0 || mode == 3) { a_hi = (1024u >> 4) | (1u << 14) | (2u << 29); a_lo0 = 0; b_hi = a_hi; }
    else if (mode == 1) { a_hi = ((uint32_t)(sbo >> 4) & 0x3FFF) | (1u << 14); a_lo0 = (((uint32_t)lbo >> 4) & 0x3FFF) << 16; b_hi = (1024u >> 4) | (1u << 14) | (2u << 29); }
    else { a_hi = (512u >> 4) | (1u << 14) | (4u << 29); a_lo0 = 0; b_hi = a_hi; }
    const uint32_t lead = elect_one() ? 1u : 0u;
    __shared__ uint64_t bars[8];
    if (ld_mode < 0) {   // emulate the real kernel's k-step protocol: `g` MMAs, commit, then wait on a barrier that is already complete
      if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
      __syncwarp();
      const int g = -ld_mode;
      uint32_t phase = 0; int st = 0;
      // pre-arrive all so that waits complete immediately (like TMA data that already landed)
      if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&bars[i]);
      __syncwarp();
      const long long t0 = clock64();
      const int flags = lbo;   // bit0 wait, bit1 fence, bit2 elect per k-step, bit3 commit
      for (int r = 0; r < reps * 4 / g; ++r) {
        if (flags & 1) mbar_wait(&bars[st], phase, 50);
        if (flags & 2) tc_fence_after();
        const uint32_t ld2 = (flags & 4) ? (elect_one() ? 1u : 0u) : lead;
        const uint32_t a0 = (sA >> 4) + (r & 7) * 8, b0 = (sB >> 4) + (r & 7) * 8;
        for (int k = 0; k < g; ++k) umma_bf16_if32(ld2, tmem, a_lo0 | (a0 + 2 * (k & 3)), a_hi, b0 + 2 * (k & 3), b_hi, idesc, (r | k) ? 1u : 0u);
        if (flags & 8) umma_commit_if(ld2, &bars[st]);     // arrives when these MMAs retire -> next phase of this barrier
        __syncwarp();
        if (++st == 8) { st = 0; phase ^= 1; }
      }
      const long long t1 = clock64();
      umma_commit_if(lead, &bar);
      __syncwarp();
      mbar_wait(&bar, 0, 99);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[blockIdx.x * 4] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; done = 1; }
      goto finish;
    }
    const long long t0 = clock64();
    if (mode == 3) {
      // accumulator alternation: consecutive MMAs go to `sbo` different accumulators (128 columns apart), switching every `lbo` MMAs
      // (lbo = log2 of the burst length, sbo = number of accumulators, a power of two: shifts and masks only)
      const int bsh = lbo, amask = sbo - 1;
      for (int r = 0; r < reps; ++r) {
        const uint32_t a0 = (sA >> 4) + (r & 7) * 8, b0 = (sB >> 4) + (r & 7) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = r * 4 + k;
          umma_bf16_if32(lead, tmem + (uint32_t)((((i >> bsh) & amask)) << 7), a_lo0 | (a0 + 2 * k), a_hi, b0 + 2 * k, b_hi, idesc, r >= 16 ? 1u : 0u);
        }
      }
    } else
    for (int r = 0; r < reps; ++r) {
      const uint32_t a0 = (sA >> 4) + (r & 7) * 8, b0 = (sB >> 4) + (r & 7) * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_if32(lead, tmem, a_lo0 | (a0 + 2 * k), a_hi, b0 + 2 * k, b_hi, idesc, (r | k) ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit_if(lead, &bar);
    __syncwarp();
    mbar_wait(&bar, 0, 99);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; done = 1; }
  } else if (warp >= 4 && ld_mode == 64) {
    // does a saturated MMA issuer (warp 0, scheduler 0) slow down ALU work of the warps that share its scheduler?
    // every warp runs the same dependent FMA chain until the issuer is done; out[8 + warp] = iterations completed
    float acc = (float)threadIdx.x;
    unsigned long long n = 0;
    while (!done) {
#pragma unroll
      for (int i = 0; i < 64; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
      ++n;
    }
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 32 + 8 + warp] = n; if (acc == 123.0f) out[0] = 1; }
  } else if (warp >= 4 && ld_mode) {
    // epilogue-like TMEM readers on accumulator columns 256.. (not written by the MMAs)
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    float acc = 0.f;
    unsigned long long n = 0;
    while (!done) {
      if (ld_mode == 16) { float v[16]; tmem_ld16(taddr + (n & 7) * 16, v); tmem_ld_wait(); acc += v[0] + v[15]; }
      else { float v[32]; tmem_ld32(taddr + (n & 3) * 32, v); tmem_ld_wait(); acc += v[0] + v[31]; }
      ++n;
      // a little ALU work between loads, like the gate math
      for (int i = 0; i < 64; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
    }
    if ((threadIdx.x & 31) == 0 && warp == 4) { out[blockIdx.x * 4 + 2] = n; out[blockIdx.x * 4 + 3] = (unsigned long long)__float_as_uint(acc); }
  }
finish:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory"); }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 32);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 2000;
  for (int ld : {0, 16, 32})
    for (int N : {48, 96, 192}) {
      cudaMemset(d, 0, 148 * 32);
      mma_rate<<<148, ld ? 384 : 128, 160 * 1024>>>(N, reps, 0, 0, 0, ld, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ld test: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[4];
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("concurrent tcgen05.ld.x%-2d by 8 warps  N=%3d  MMA %.1f cyc (ideal %d)   loads per warp during run: %llu  (= 1 per %.0f cyc)\n", ld, N,
             (double)h[1] / (reps * 4), N / 2 > 40 ? N / 2 : 40, h[2], h[2] ? (double)h[1] / h[2] : 0.0);
    }
  {
    unsigned long long* d2;
    cudaMalloc(&d2, 148 * 32 * 8);
    for (int N : {32, 96, 256}) {
      cudaMemset(d2, 0, 148 * 32 * 8);
      mma_rate<<<1, 384, 160 * 1024>>>(N, reps * 8, 0, 0, 0, 64, d2);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("scheduler test: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[32];
      cudaMemcpy(h, d2, 256, cudaMemcpyDeviceToHost);
      printf("issuer on scheduler 0, N=%3d, %.1f cyc/MMA: FMA-chain iterations of warps 4..11 (scheduler = warp %% 4):", N, (double)h[1] / (reps * 8 * 4));
      for (int w = 4; w < 12; ++w) printf(" %llu", h[8 + w]);
      printf("\n");
    }
  }
  for (int N : {32, 48, 96})
    for (int nacc : {1, 2, 4})
      for (int bsh : {0, 1, 3, 4}) {
        const int burst = 1 << bsh;
        if (nacc == 1 && burst > 1) continue;
        mma_rate<<<148, 128, 160 * 1024>>>(N, reps, 3, bsh, nacc, 0, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("acc test: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("accumulators=%d switch every %2d MMAs  N=%3d  issue %.1f cyc/MMA  complete %.1f cyc/MMA\n", nacc, burst, N, (double)h[0] / (reps * 4), (double)h[1] / (reps * 4));
      }
  struct Case { int mode, lbo, sbo; const char* name; } cases[] = {
      {0, 0, 0, "A SW128 / B SW128"}, {2, 0, 0, "A SW64 / B SW64"}, {1, 2880, 160, "A no-swizzle LBO=2880 SBO=160 (halo 18x10)"},
      {1, 2048, 128, "A no-swizzle LBO=2048 SBO=128 (per-tap 16x8)"}, {1, 16, 208, "A no-swizzle LBO=16 SBO=208 (stem)"},
      {1, 128, 256, "A no-swizzle LBO=128 SBO=256 (canonical packed)"}};
  for (auto& c : cases)
    for (int N : {32, 48, 96, 192, 256})
      for (int grid : {1, 148}) {
        mma_rate<<<grid, 128, 160 * 1024>>>(N, reps, c.mode, c.lbo, c.sbo, 0, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s N=%d: %s\n", c.name, N, cudaGetErrorString(e)); return 1; }
        unsigned long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-48s N=%3d grid=%3d  issue %.1f cyc/MMA   complete %.1f cyc/MMA  (ideal %d)\n", c.name, N, grid, (double)h[0] / (reps * 4),
               (double)h[1] / (reps * 4), N / 2);
      }
  for (int flags : {0, 2, 4, 8, 9, 11, 15})
   for (int g : {2, 6})
    for (int N : {48}) {
      cudaMemset(d, 0, 148 * 32);
      mma_rate<<<148, 128, 160 * 1024>>>(N, reps, 0, flags, 0, -g, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kstep test: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[4];
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("k-step protocol flags=%2d (1 wait,2 fence,4 elect,8 commit) %2d MMAs/k-step N=%3d: %.1f cyc/MMA -> %.0f cyc per k-step (pipe %d/MMA)\n", flags, g, N,
             (double)h[1] / (reps * 4), (double)h[1] / (reps * 4) * g, N / 2 > 40 ? N / 2 : 40);
    }
  return 0;
}
