// tcgen05 implicit-GEMM convolution for sm_100a.
//
// One persistent CTA per SM. Per output tile (8 x 16 positions of one image = UMMA M = 128) the
// K loop walks (tap, channel chunk): the A operand is a TMA *tiled* box of the NHWC input shifted by the
// tap offset (zero padding = TMA out-of-bounds fill; stride-2 = TMA element strides) -- 64-channel chunks
// land as 128 B rows (SWIZZLE_128B), a trailing 32-channel chunk as 64 B rows (SWIZZLE_64B); the B operand
// of a whole pipeline stage is ONE linear cp.async.bulk of the pre-swizzled weight image. Both feed
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) accumulating into one of two TMEM
// accumulator stages; four epilogue warps drain the other stage (tcgen05.ld), apply
// bias + ELU/ReLU x sigmoid gating (reference models/networks/utils.py:25-33) or the linear
// epilogue of the attention GEMMs, and store NHWC.
//
//   warp 0 : TMA producer (one lane)        warp 1 : MMA issuer (one lane)
//   warp 2 : TMEM allocator                 warps 4-7 : epilogue (TMEM lane quadrant = warp % 4)
#include "se_common.cuh"
#include "se_conv_tc.h"

#include <stdlib.h>

#include <vector>

namespace se {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("se_conv_tc: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
  __trap();
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// linear global -> shared bulk copy (bytes % 16 == 0), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one lane of a fully converged warp (the warp stays converged: the compiler keeps addresses / descriptors
// in uniform registers instead of broadcasting them lane by lane)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M=128, N from idesc, K=16.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// must be executed before the registers written by tmem_ld16 are read
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write cnt (<= 16) consecutive channels of one pixel; static register indexing only (no local memory)
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* o, const float (&r)[16], int cnt, bool al8, bool al4) {
  if (cnt == 16 && al8) {
    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
    *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]), pack_bf16x2(r[14], r[15]));
  } else if (al4) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * j + 4 <= cnt) *reinterpret_cast<uint2*>(o + 4 * j) = make_uint2(pack_bf16x2(r[4 * j], r[4 * j + 1]), pack_bf16x2(r[4 * j + 2], r[4 * j + 3]));
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i >= (cnt & ~3) && i < cnt) o[i] = __float2bfloat16(r[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < cnt) o[i] = __float2bfloat16(r[i]);
  }
}
__device__ __forceinline__ void store_row_f32(float* o, const float (&r)[16], int cnt, bool al4) {
  if (cnt == 16 && al4) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < cnt) o[i] = r[i];
  }
}

// K-major operand tiles (cute::UMMA::SmemDescriptor): rows of 128 B (SWIZZLE_128B, 8-row atoms 1024 B apart)
// or rows of 64 B (SWIZZLE_64B, atoms 512 B apart). Only the 14-bit start-address field changes per MMA.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, bool sw128) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);               // start address        bits [0,14)
  d |= (uint64_t)((sw128 ? 1024 : 512) >> 4) << 32;           // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                                     // descriptor version 1 (sm_100)
  d |= (uint64_t)(sw128 ? 2 : 4) << 61;                       // layout type: SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// ------------------------------------------------------------------------------------------ kernel
constexpr int A64_BYTES = TILE_M * 128;              // 64-channel A unit
constexpr int A32_BYTES = TILE_M * 64;               // 32-channel A unit
constexpr int NUM_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;                      // TMEM columns between the two accumulator stages
constexpr int MAX_STAGES = 8;

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmA32, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A64 x r64 | A32 x r32 | B image] then barriers, tmem ptr, bias
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.r64 * A64_BYTES + p.r32 * A32_BYTES;
  const int b64_bytes = p.NT * 128, b32_bytes = p.NT * 64;
  const int b_bytes = p.r64 * b64_bytes + p.r32 * b32_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* tail = smem + (size_t)p.num_stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_smem + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA64)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA32)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias (all N tiles) into shared memory; zero when absent
  for (int i = threadIdx.x; i < p.n_tiles * p.NT + 32; i += NUM_THREADS) bias_s[i] = (p.bias != nullptr && i < p.Cout) ? p.bias[i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.N * tiles_per_img * p.n_tiles;
  const int ksteps = p.ksteps;

  if (warp == 0) {
    // ==================================================================== TMA producer (warp converged, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    long long t_wait = 0, t_begin = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      int rest = tile / p.n_tiles;
      const int tx = rest % p.tiles_x;
      rest /= p.tiles_x;
      const int ty = rest % p.tiles_y;
      const int img = rest / p.tiles_y;
      const int x0 = tx * TILE_W * p.stride, y0 = ty * TILE_H * p.stride;
      const uint8_t* wsrc = p.w + (size_t)img * p.w_img_bytes + (size_t)nt * ksteps * b_bytes;
      for (int ks = 0; ks < ksteps; ++ks) {
        const long long tw = clock64();
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        t_wait += clock64() - tw;
        uint8_t* sA = smem + (size_t)stage * stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          bulk_load_1d(sA + a_bytes, wsrc + (size_t)ks * b_bytes, (uint32_t)b_bytes, &full_bar[stage]);
          for (int j = 0; j < p.r64; ++j) {
            const int u = ks * p.r64 + j;
            const int t = u / p.n64, chunk = u - t * p.n64;
            tma_load_4d(sA + j * A64_BYTES, &tmA64, &full_bar[stage], chunk * 64, x0 + p.dx[t], y0 + p.dy[t], img);
          }
          for (int j = 0; j < p.r32; ++j) {
            const int t = ks * p.r32 + j;     // n32 == 1: one 32-wide unit per tap
            tma_load_4d(sA + p.r64 * A64_BYTES + j * A32_BYTES, &tmA32, &full_bar[stage], p.n64 * 64, x0 + p.dx[t], y0 + p.dy[t], img);
          }
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 0] = t_wait; p.dbg[blockIdx.x * 8 + 1] = clock64() - t_begin; }
  } else if (warp == 1) {
    // ==================================================================== MMA issuer (warp converged, one lane issues)
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    long long t_wfull = 0, t_wtmem = 0, t_begin = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      long long tw = clock64();
      mbar_wait(&tmem_empty[as], aphase ^ 1, 2);
      t_wtmem += clock64() - tw;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
      for (int ks = 0; ks < ksteps; ++ks) {
        tw = clock64();
        mbar_wait(&full_bar[stage], phase, 3);
        t_wfull += clock64() - tw;
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sB = sA + a_bytes;
        if (elect_one()) {
          uint32_t acc = ks ? 1u : 0u;
          // 64-channel units: four K=16 MMAs each, +32 B (= +2 in the descriptor) per step
          uint64_t ad = make_kmajor_desc(sA, true), bd = make_kmajor_desc(sB, true);
          for (int j = 0; j < p.r64; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tmem_d, ad + 2 * k, bd + 2 * k, idesc, acc);
              acc = 1u;
            }
            ad += A64_BYTES >> 4;
            bd += (uint32_t)b64_bytes >> 4;
          }
          // trailing 32-channel units: two K=16 MMAs each
          ad = make_kmajor_desc(sA + p.r64 * A64_BYTES, false);
          bd = make_kmajor_desc(sB + p.r64 * b64_bytes, false);
          for (int j = 0; j < p.r32; ++j) {
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            umma_bf16(tmem_d, ad + 2, bd + 2, idesc, 1u);
            acc = 1u;
            ad += A32_BYTES >> 4;
            bd += (uint32_t)b32_bytes >> 4;
          }
          umma_commit(&empty_bar[stage]);                       // smem slot free once these MMAs retire
          if (ks == ksteps - 1) umma_commit(&tmem_full[as]);    // accumulator complete
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 2] = t_wfull; p.dbg[blockIdx.x * 8 + 3] = t_wtmem; p.dbg[blockIdx.x * 8 + 4] = clock64() - t_begin; }
  } else if (warp >= 4) {
    // ==================================================================== epilogue
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;             // tile row == TMEM lane == output position in tile
    const int ry = row / TILE_W, rx = row % TILE_W;
    int iter = 0;
    long long t_wacc = 0, t_begin = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int nt = tile % p.n_tiles;
      int rest = tile / p.n_tiles;
      const int tx = rest % p.tiles_x;
      rest /= p.tiles_x;
      const int ty = rest % p.tiles_y;
      const int img = rest / p.tiles_y;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      const long long tw = clock64();
      mbar_wait(&tmem_full[as], aphase, 4);
      t_wacc += clock64() - tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
      const int py = ty * TILE_H + ry, px = tx * TILE_W + rx;
      const bool valid = (py < p.Ho) && (px < p.Wo);
      const int oy = py * p.osy + p.ooy, ox = px * p.osx + p.oox;
      const size_t opix = ((size_t)img * p.Hout + oy) * p.Wout + ox;

      if (p.epi == EPI_LINEAR) {
        const int n0 = nt * p.NT;
        const float* cs = p.colscale ? p.colscale + (size_t)img * p.Cout : nullptr;
        const bool al4 = ((p.ldo | p.choff) & 3) == 0, al8 = ((p.ldo | p.choff) & 7) == 0;
        for (int c0 = 0; c0 < p.NT; c0 += 16) {
          const int cb = n0 + c0;
          if (cb >= p.Cout) break;
          float v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (valid) {
            const int cnt = min(16, p.Cout - cb);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float sc = p.scale;
              if (cs != nullptr && i < cnt) sc *= __ldg(cs + cb + i);
              v[i] = (v[i] + bias_s[cb + i]) * sc;
            }
            if (p.out_dt == DT_F32) store_row_f32(reinterpret_cast<float*>(p.y) + opix * p.ldo + p.choff + cb, v, cnt, al4);
            else store_row_bf16(reinterpret_cast<__nv_bfloat16*>(p.y) + opix * p.ldo + p.choff + cb, v, cnt, al8, al4);
          }
        }
      } else {
        // gated: feature column c pairs with gate column c + half; both live in this thread's lane
        const int half = p.Cout >> 1;
        const bool is_elu = (p.epi == EPI_GATE_ELU);
        const bool al4 = ((p.ldo | p.choff) & 3) == 0, al8 = ((p.ldo | p.choff) & 7) == 0;
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.y) + opix * p.ldo + p.choff;
        for (int c0 = 0; c0 < half; c0 += 16) {
          float f[16], g[16];
          tmem_ld16(taddr + c0, f);
          tmem_ld16(taddr + half + c0, g);
          tmem_ld_wait();
          if (valid) {
            const int cnt = min(16, half - c0);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float fv = f[i] + bias_s[c0 + i], gv = g[i] + bias_s[half + c0 + i];
              const float a = is_elu ? (fv > 0.0f ? fv : ex2_approx(fv * 1.4426950408889634f) - 1.0f) : fmaxf(fv, 0.0f);
              f[i] = a * fmaf(0.5f, tanh_approx(0.5f * gv), 0.5f);     // a * sigmoid(g)
            }
            store_row_bf16(o + c0, f, cnt, al8, al4);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
    if (p.dbg && threadIdx.x == 128) { p.dbg[blockIdx.x * 8 + 5] = t_wacc; p.dbg[blockIdx.x * 8 + 6] = clock64() - t_begin; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int g_num_sms = 0;
static int g_smem_optin = 0;

static int tc_smem_budget() { return 200 * 1024; }

void tc_choose_stage(TcWeights* w) {
  static int cap_kb = getenv("SE_TC_STAGE_KB") ? atoi(getenv("SE_TC_STAGE_KB")) : 48;
  const int u64 = A64_BYTES + w->NT * 128, u32 = A32_BYTES + w->NT * 64;
  if (w->n64 > 0 && w->n32 > 0) {
    // mixed taps (e.g. 96 channels = 64 + 32): a stage holds whole taps
    int best = 1;
    for (int t = 1; t <= 4 && t <= w->ntaps; ++t)
      if (w->ntaps % t == 0 && t * (w->n64 * u64 + u32) <= cap_kb * 1024) best = t;
    w->r64 = best * w->n64;
    w->r32 = best;
  } else if (w->n64 > 0) {
    const int total = w->ntaps * w->n64;
    int best = 1;
    for (int k = 1; k <= 8 && k <= total; ++k)
      if (total % k == 0 && k * u64 <= cap_kb * 1024) best = k;
    w->r64 = best;
    w->r32 = 0;
  } else {
    const int total = w->ntaps * w->n32;
    int best = 1;
    for (int k = 1; k <= 8 && k <= total; ++k)
      if (total % k == 0 && k * u32 <= cap_kb * 1024) best = k;
    w->r64 = 0;
    w->r32 = best;
  }
}

int tc_plan(const ConvParams& c, const TcWeights& w, TcParams* out, int* smem_bytes) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  SE_REQUIRE(c.in_dt == DT_BF16, "tcgen05 path reads bf16 activations");
  SE_REQUIRE(c.ldx % 8 == 0, "input pixel pitch must be a multiple of 8 elements (16 B TMA stride)");
  SE_REQUIRE((reinterpret_cast<uintptr_t>(c.x) & 15) == 0, "input base must be 16 B aligned");
  SE_REQUIRE((reinterpret_cast<uintptr_t>(w.data) & 15) == 0 && w.img_bytes % 16 == 0, "weight image must be 16 B aligned");
  SE_REQUIRE(c.ntaps <= MAX_TAPS && c.ntaps == w.ntaps, "tap count mismatch");
  SE_REQUIRE(w.NT % 16 == 0 && w.NT >= 16 && w.NT <= 256, "NT");
  SE_REQUIRE(w.n32 <= 1 && (w.n64 > 0 || w.n32 > 0), "chunking");
  SE_REQUIRE(w.n64 * 64 + w.n32 * 32 >= c.Ci, "weights do not cover Ci");
  SE_REQUIRE((w.n64 == 0 || (w.r64 > 0 && (w.ntaps * w.n64) % w.r64 == 0)) && (w.n32 == 0 || (w.r32 > 0 && w.ntaps % w.r32 == 0)), "stage grouping");
  SE_REQUIRE(w.n64 == 0 || w.n32 == 0 || w.ntaps * w.n64 / w.r64 == w.ntaps / w.r32, "mixed stages must hold whole taps");
  SE_REQUIRE(c.stride >= 1 && c.stride <= 2, "stride");
  p.N = c.N; p.Ho = c.Ho; p.Wo = c.Wo;
  p.tiles_x = (c.Wo + TILE_W - 1) / TILE_W;
  p.tiles_y = (c.Ho + TILE_H - 1) / TILE_H;
  p.n_tiles = w.n_tiles;
  p.stride = c.stride;
  p.ntaps = c.ntaps;
  memcpy(p.dy, c.dy, sizeof(p.dy));
  memcpy(p.dx, c.dx, sizeof(p.dx));
  p.n64 = w.n64; p.n32 = w.n32; p.r64 = w.n64 ? w.r64 : 0; p.r32 = w.n32 ? w.r32 : 0; p.NT = w.NT;
  p.ksteps = tc_ksteps(w);
  p.w_img_bytes = w.img_bytes;
  p.w = reinterpret_cast<const uint8_t*>(w.data);
  p.bias = c.bias; p.Cout = c.Cout;
  p.y = c.y; p.out_dt = c.out_dt; p.Hout = c.Hout; p.Wout = c.Wout; p.ldo = c.ldo; p.choff = c.choff;
  p.osy = c.osy; p.ooy = c.ooy; p.osx = c.osx; p.oox = c.oox;
  p.epi = c.epi; p.scale = c.scale; p.colscale = c.colscale;
  if (c.epi != EPI_LINEAR) {
    SE_REQUIRE(w.n_tiles == 1 && c.Cout % 2 == 0 && c.out_dt == DT_BF16, "gated epilogue needs one N tile, even Cout, bf16 out");
  }
  const int stage_bytes = p.r64 * (A64_BYTES + w.NT * 128) + p.r32 * (A32_BYTES + w.NT * 64);
  int stages = tc_smem_budget() / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SE_REQUIRE(stages >= 2, "pipeline needs >= 2 stages");
  p.num_stages = stages;
  *smem_bytes = 1024 + stages * stage_bytes + (2 * MAX_STAGES + 4) * 8 + 16 + (p.n_tiles * p.NT + 32) * 4 + 64;
  *out = p;
  return 0;
}

static int encode_act_map(EncodeTiledFn enc, CUtensorMap* tm, const ConvParams& c, int inner, CUtensorMapSwizzle swz) {
  // activations: (C, W, H, N), box (inner, 16*s, 8*s, 1) walked with element strides (1, s, s, 1)
  const long long row_pitch = c.x_row_pitch ? c.x_row_pitch : (long long)c.Wi * c.ldx;
  const long long img_pitch = c.x_img_pitch ? c.x_img_pitch : (long long)c.Hi * row_pitch;
  cuuint64_t dims[4] = {(cuuint64_t)c.Ci, (cuuint64_t)c.Wi, (cuuint64_t)c.Hi, (cuuint64_t)c.N};
  cuuint64_t strides[3] = {(cuuint64_t)c.ldx * 2, (cuuint64_t)row_pitch * 2, (cuuint64_t)img_pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)(TILE_W * c.stride), (cuuint32_t)(TILE_H * c.stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(c.x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed, CUresult=" + std::to_string((int)r));
  return 0;
}

int tc_launch(const ConvParams& c, const TcWeights& w, cudaStream_t stream) {
  TcParams p;
  int smem_bytes = 0;
  int rc = tc_plan(c, w, &p, &smem_bytes);
  if (rc) return rc;
  EncodeTiledFn enc = get_encode_fn();
  SE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  if (!g_num_sms) {
    int dev = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    SE_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    SE_CUDA_OK(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SE_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
  }
  SE_REQUIRE(smem_bytes <= g_smem_optin, "shared memory plan exceeds the opt-in limit");

  CUtensorMap tmA64, tmA32;
  rc = encode_act_map(enc, &tmA64, c, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_act_map(enc, &tmA32, c, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;

  const int total_tiles = p.N * p.tiles_x * p.tiles_y * p.n_tiles;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  static const bool dbg_on = getenv("SE_TC_DEBUG") != nullptr;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on) {
    if (!dbg_buf) SE_CUDA_OK(cudaMalloc(&dbg_buf, 8 * 8 * 1024));
    SE_CUDA_OK(cudaMemsetAsync(dbg_buf, 0, 8 * 8 * 1024, stream));
    p.dbg = dbg_buf;
  }
  conv_tc_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tmA64, tmA32, p);
  SE_CUDA_OK(cudaGetLastError());
  if (dbg_on) {
    SE_CUDA_OK(cudaStreamSynchronize(stream));
    std::vector<unsigned long long> h(8 * grid);
    SE_CUDA_OK(cudaMemcpy(h.data(), dbg_buf, h.size() * 8, cudaMemcpyDeviceToHost));
    double a[8] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 8; ++k) a[k] += (double)h[b * 8 + k] / grid;
    fprintf(stderr,
            "[tc] N=%d %dx%d Ci=%d s=%d taps=%d NT=%d n64=%d n32=%d r64=%d r32=%d stages=%d tiles=%d grid=%d | prod wait_empty %.0f/%.0f | mma wait_full %.0f wait_tmem %.0f /%.0f | epi wait_acc %.0f/%.0f (cycles, mean per CTA)\n",
            c.N, c.Ho, c.Wo, c.Ci, c.stride, c.ntaps, p.NT, p.n64, p.n32, p.r64, p.r32, p.num_stages, total_tiles, grid, a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
  }
  return 0;
}

}  // namespace se
