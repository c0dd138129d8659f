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
#include "se_tc_device.cuh"

#include <stdlib.h>

#include <vector>

namespace se {

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
constexpr int NUM_THREADS = TC_NUM_THREADS;
constexpr int TMEM_COLS = TC_TMEM_COLS;
constexpr int ACC_STRIDE = TC_ACC_STRIDE;
constexpr int MAX_STAGES = TC_MAX_STAGES;

// R64 / R32: compile-time copies of p.r64 / p.r32 (R64 < 0: run-time values). Compile-time trip counts let the MMA
// issue sequence of a k-step unroll completely (a rolled loop is issue-latency bound at ~110 cycles per MMA,
// tools/bench/mma_rate.cu).
template <int R64, int R32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmA32, const TcParams p) {
  const int r64 = R64 >= 0 ? R64 : p.r64, r32 = R64 >= 0 ? R32 : p.r32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A64 x r64 | A32 x r32 | B image] then barriers, tmem ptr, bias
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = r64 * A64_BYTES + r32 * A32_BYTES;
  const int b64_bytes = p.NT * 128, b32_bytes = p.NT * 64;
  const int b_bytes = r64 * b64_bytes + r32 * b32_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* tail = smem + (size_t)p.num_stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~uintptr_t(15));   // 16 B aligned: read with ld.shared.v4

  // warp-uniform by construction (a shuffle from lane 0): lets the compiler keep role-dependent values - the tile
  // parity of the second MMA issuer, ring indices, descriptors - on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
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
      mbar_init(&tmem_empty[i], TC_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias (all N tiles) into shared memory; zero when absent
  const int cst_n = (p.e.has_bias ? p.n_tiles * p.NT : 0) + 32;   // attention S has up to 32 N tiles and no bias: nothing to stage
  epi_fill_constants(bias_s, cst_n, p.bias, p.e, threadIdx.x, NUM_THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // uniform copy (feeds MMA / tcgen05.ld addresses)

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
        const long long tw = p.dbg ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        if (p.dbg) t_wait += clock64() - tw;
        uint8_t* sA = smem + (size_t)stage * stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          bulk_load_1d(sA + a_bytes, wsrc + (size_t)ks * b_bytes, (uint32_t)b_bytes, &full_bar[stage]);
          for (int j = 0; j < r64; ++j) {
            const int u = ks * r64 + j;
            const int t = u / p.n64, chunk = u - t * p.n64;
            tma_load_4d(sA + j * A64_BYTES, &tmA64, &full_bar[stage], chunk * 64, x0 + p.dx[t], y0 + p.dy[t], img);
          }
          for (int j = 0; j < r32; ++j) {
            const int t = ks * r32 + j;     // n32 == 1: one 32-wide unit per tap
            tma_load_4d(sA + r64 * A64_BYTES + j * A32_BYTES, &tmA32, &full_bar[stage], p.n64 * 64, x0 + p.dx[t], y0 + p.dy[t], img);
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
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // shared-window address as a plain integer: uniform
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    long long t_wfull = 0, t_wtmem = 0, t_begin = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      long long tw = p.dbg ? clock64() : 0;
      mbar_wait(&tmem_empty[as], aphase ^ 1, 2);
      if (p.dbg) t_wtmem += clock64() - tw;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
      for (int ks = 0; ks < ksteps; ++ks) {
        tw = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase, 3);
        if (p.dbg) t_wfull += clock64() - tw;
        tc_fence_after();
        const uint32_t sA = smem_base + (uint32_t)stage * stage_bytes;
        const uint32_t sB = sA + a_bytes;
        {
          const uint32_t lead = elect_one() ? 1u : 0u;   // predicated issue: the warp stays converged
          uint32_t acc = ks ? 1u : 0u;
          const uint32_t hi128 = (1024u >> 4) | (1u << 14) | (2u << 29), hi64 = (512u >> 4) | (1u << 14) | (4u << 29);
          // 64-channel units: four K=16 MMAs each, +32 B (= +2 in the descriptor) per step
#pragma unroll
          for (int j = 0; j < (R64 >= 0 ? R64 : 16); ++j) {
            if (j < r64) {
              const uint32_t a0 = (sA + j * A64_BYTES) >> 4, b0 = (sB + j * b64_bytes) >> 4;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_if32(lead, tmem_d, a0 + 2 * k, hi128, b0 + 2 * k, hi128, idesc, acc);
                acc = 1u;
              }
            }
          }
          // trailing 32-channel units: two K=16 MMAs each
#pragma unroll
          for (int j = 0; j < (R64 >= 0 ? R32 : 16); ++j) {
            if (j < r32) {
              const uint32_t a0 = (sA + r64 * A64_BYTES + j * A32_BYTES) >> 4, b0 = (sB + r64 * b64_bytes + j * b32_bytes) >> 4;
              umma_bf16_if32(lead, tmem_d, a0, hi64, b0, hi64, idesc, acc);
              umma_bf16_if32(lead, tmem_d, a0 + 2, hi64, b0 + 2, hi64, idesc, 1u);
              acc = 1u;
            }
          }
          umma_commit_if(lead, &empty_bar[stage]);                       // smem slot free once these MMAs retire
          if (ks == ksteps - 1) umma_commit_if(lead, &tmem_full[as]);    // accumulator complete
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 2] = t_wfull; p.dbg[blockIdx.x * 8 + 3] = t_wtmem; p.dbg[blockIdx.x * 8 + 4] = clock64() - t_begin; }
  } else if (warp >= 4) {
    // ==================================================================== epilogue
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;            // which share of the 16-column chunks this warp drains
    const bool fast_epi = epi_fast_ok(p.e);
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
      const long long tw = p.dbg ? clock64() : 0;
      mbar_wait(&tmem_full[as], aphase, 4);
      if (p.dbg) t_wacc += clock64() - tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
      const int py = ty * TILE_H + ry, px = tx * TILE_W + rx;
      const bool valid = (py < p.Ho) && (px < p.Wo);
      const int oy = py * p.e.osy + p.e.ooy, ox = px * p.e.osx + p.e.oox;
      if (fast_epi) tc_epilogue_tile<true>(p.e, bias_s, cst_n, taddr, img, nt, valid, oy, ox, grp);
      else tc_epilogue_tile<false>(p.e, bias_s, cst_n, taddr, img, nt, valid, oy, ox, grp);
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

void fill_epi(const ConvParams& c, int NT, EpiParams* e) {
  e->y = c.y; e->out_dt = c.out_dt; e->out_c8 = c.out_c8;
  e->Hout = c.Hout; e->Wout = c.Wout; e->ldo = c.ldo; e->choff = c.choff;
  e->osy = c.osy; e->ooy = c.ooy; e->osx = c.osx; e->oox = c.oox;
  e->epi = c.epi; e->scale = c.scale; e->colscale = c.colscale;
  e->Cout = c.Cout; e->NT = NT;
  e->has_bias = c.bias != nullptr ? 1 : 0;
  e->blk_split = c.out_blk_split > 0 ? c.out_blk_split : (1 << 20);
  e->blk_jump = c.out_blk_split > 0 ? c.out_blk_jump : 0;
  e->par_stride = c.out_par_stride > 0 ? c.out_par_stride : (c.ldo >> 2);
  e->nsplit = c.f16x2 ? 2 : 1;
  e->split_stride = (int)c.out_split_stride;
  e->goff = (c.epi == EPI_LINEAR) ? 0 : gated_goff(c.Cout);
}
// the fast epilogue addresses the output in 32-bit units of 16 B
static inline bool epi_out_fits_u32(const ConvParams& c) {
  const double units = c.out_c8 ? (double)c.N * c.ldo * c.Hout * c.Wout : (double)c.N * c.Hout * c.Wout * c.ldo / 8.0;
  return units < 4294967296.0 && (reinterpret_cast<uintptr_t>(c.y) & 15) == 0;
}
bool epi_addressable(const ConvParams& c) { return epi_out_fits_u32(c); }

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
  p.bias = c.bias;
  fill_epi(c, w.NT, &p.e);
  SE_REQUIRE(c.out_dt != DT_BF16 || epi_addressable(c), "output tensor too large / misaligned for 32-bit block addressing");
  SE_REQUIRE(!c.out_c8 || (c.out_dt == DT_BF16 && c.choff % 8 == 0), "C8 output must be bf16 with a channel offset multiple of 8");
  SE_REQUIRE(c.out_c8 != 2, "space-to-depth output is written by the channel-blocked kernel only");
  if (c.epi != EPI_LINEAR) {
    SE_REQUIRE(w.n_tiles == 1 && c.Cout % 2 == 0 && c.out_dt == DT_BF16, "gated epilogue needs one N tile, even Cout, bf16 out");
  }
  const int stage_bytes = p.r64 * (A64_BYTES + w.NT * 128) + p.r32 * (A32_BYTES + w.NT * 64);
  int stages = tc_smem_budget() / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SE_REQUIRE(stages >= 2, "pipeline needs >= 2 stages");
  p.num_stages = stages;
  *smem_bytes = 1024 + stages * stage_bytes + (2 * MAX_STAGES + 4) * 8 + 16 + 3 * ((c.bias ? p.n_tiles * p.NT : 0) + 32) * 4 + 64;
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

// k-step structures used by the generator: stride-2 layers (0,3) (1,0), attention S (1,1) and AV (1,0)
#define TC_SPECIALISATIONS(X) X(0, 3) X(1, 0) X(1, 1)

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
#define X(a, b) SE_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<a, b>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    TC_SPECIALISATIONS(X)
#undef X
    SE_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<-1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
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
  bool launched = false;
#define X(a, b)                                                                              \
  if (!launched && p.r64 == a && p.r32 == b) {                                                \
    conv_tc_kernel<a, b><<<grid, NUM_THREADS, smem_bytes, stream>>>(tmA64, tmA32, p);        \
    launched = true;                                                                          \
  }
  TC_SPECIALISATIONS(X)
#undef X
  if (!launched) conv_tc_kernel<-1, 0><<<grid, NUM_THREADS, smem_bytes, stream>>>(tmA64, tmA32, p);
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
