// Glue kernels of the fp32-on-tensor-cores mode (SE_PREC_FP32_TC): activations are stored "split half" (DT_F16X2, se_common.cuh):
// value * 64 (kSplitActScale, se_common.cuh) = hi + lo with hi = fp16(64 v), lo = fp16(64 v - hi), in channel-blocked tensors whose lo blocks follow the hi blocks
// ([N][2*CB][H][W][8]; space-to-depth: [N][4 parities][2][CB][H/2][W/2][8]). Everything here is the split-half twin of a
// kernel in se_misc.cu: input packing, heads, global pooling, layout conversion (reference call sites are cited there).
#include <cuda_fp16.h>

#include "se_common.cuh"
#include "se_misc.h"

namespace se {

static inline int cdiv_s(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ void split8(const float (&v)[8], uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const float s0 = fminf(fmaxf(v[k] * kSplitActScale, -kSplitActMax), kSplitActMax), s1 = fminf(fmaxf(v[k + 1] * kSplitActScale, -kSplitActMax), kSplitActMax);
    const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
    const __half l0 = __float2half_rn(s0 - __half2float(h0)), l1 = __float2half_rn(s1 - __half2float(h1));
    h[k >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[k >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __half22float2(h[k]), b = __half22float2(l[k]);
    v[2 * k] = (a.x + b.x) * kSplitActInv;
    v[2 * k + 1] = (a.y + b.y) * kSplitActInv;
  }
}

// ------------------------------------------------------------------------------------------ pack8 (se_misc.cu: pack8_kernel)
// out: [B][2 blocks (hi, lo)][H][Wp][8]
__global__ void pack8_split_kernel(const float* __restrict__ img, const float* __restrict__ sketch, const float* __restrict__ mask, uint4* __restrict__ out,
                                   int B, int H, int W, int Wp, int padl, int img_mode, float sketch_scale, int write_mask) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (b, y, xp)
  const long long HW = (long long)H * W, plane = (long long)H * Wp;
  if (j >= (long long)B * plane) return;
  const int xp = (int)(j % Wp);
  const long long by = j / Wp;
  const long long b = by / H;
  const long long o = (b * 2) * plane + (by % H) * Wp + xp;
  const int x = xp - padl;
  if (x < 0 || x >= W) {
    out[o] = make_uint4(0, 0, 0, 0);
    out[o + plane] = make_uint4(0, 0, 0, 0);
    return;
  }
  const long long pix = (by % H) * W + x, i = b * HW + pix;
  const float m = mask ? mask[i] : 0.0f;
  const float a = img_mode == PACK_IMG_ONE ? 1.0f : (img_mode == PACK_IMG_ONE_MINUS_M ? 1.0f - m : m);
  float v[8];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = img[(b * 3 + c) * HW + pix] * a;
  v[3] = (sketch ? sketch[i] : 1.0f) * sketch_scale;
  v[4] = write_mask ? m : 0.0f;
  v[5] = v[6] = v[7] = 0.0f;
  split8(v, &out[o], &out[o + plane]);
}
int pack8_split(const float* img, const float* sketch, const float* mask, void* out, int B, int H, int W, int Wp, int padl, int img_mode,
                float sketch_scale, int write_mask, cudaStream_t s) {
  const long long n = (long long)B * H * Wp;
  pack8_split_kernel<<<cdiv_s(n, 256), 256, 0, s>>>(img, sketch, mask, (uint4*)out, B, H, W, Wp, padl, img_mode, sketch_scale, write_mask);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ heads (se_misc.cu: head_kernel)
// x: [B][4 blocks: hi 0-7, hi 8-15, lo 0-7, lo 8-15][H][W][8]; weights [9][12][COUT] fp32; same modes / outputs as head_kernel,
// out_pack8 (HEAD_COARSE) is the split-half packed stage-2 input [B][2][H][Wp][8]
template <int COUT>
__global__ void __launch_bounds__(128) head_split_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int B, int H,
                                                         int W, int mode, const float* __restrict__ img, const float* __restrict__ mask_bin,
                                                         const float* __restrict__ mask_soft, float* __restrict__ out_nchw, float* __restrict__ out2,
                                                         uint4* __restrict__ out_pack8, int no_mask_coarse, int Wp, int padl, long long obs, long long msbs,
                                                         unsigned char* __restrict__ out_u8) {
  __shared__ float ws[9 * 12 * COUT + COUT];
  for (int i = threadIdx.x; i < 9 * 12 * COUT; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < COUT) ws[9 * 12 * COUT + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W;
  if (i >= B * HW) return;
  const long long b = i / HW, pix = i % HW;
  const int yy = (int)(pix / W), xx = (int)(pix % W);
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = ws[9 * 12 * COUT + o];
  const uint4* p0 = x + (b * 4) * HW;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const long long q = (long long)iy * W + ix;
    float v0[8], v1[8];
    join8(p0[q], p0[2 * HW + q], v0);
    join8(p0[HW + q], p0[3 * HW + q], v1);
#pragma unroll
    for (int c = 0; c < 12; ++c) {
      const float xv = c < 8 ? v0[c] : v1[c - 8];
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] = fmaf(xv, ws[(t * 12 + c) * COUT + o], acc[o]);
    }
  }
  if (mode == HEAD_MASK) {
    const float sg = 1.0f / (1.0f + expf(-acc[0]));
    out_nchw[b * obs + pix] = sg;
    out2[i] = sg > 0.5f ? 1.0f : 0.0f;
    if (out_u8) out_u8[i] = (unsigned char)(int)(sg * 255.0f);
    return;
  }
  float t3[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) t3[o] = tanhf(acc[o]);
  if (mode == HEAD_TANH) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) out_nchw[b * obs + o * HW + pix] = t3[o];
  } else if (mode == HEAD_COARSE) {
    const float m = mask_bin[i];
    float v[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) v[o] = 0.0f;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      if (out_nchw) out_nchw[b * obs + o * HW + pix] = t3[o];
      const float xin = img[(b * 3 + o) * HW + pix] * (1.0f - m);
      v[o] = no_mask_coarse ? t3[o] : (t3[o] * m + xin * (1.0f - m));
    }
    const long long plane = (long long)H * Wp, op = (b * 2) * plane + (long long)yy * Wp + xx + padl;
    split8(v, &out_pack8[op], &out_pack8[op + plane]);
  } else {  // HEAD_FINE
    const float m = mask_soft[b * msbs + pix];
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      if (out2) out2[(b * COUT + o) * HW + pix] = t3[o];
      const float cv = t3[o] * m + img[(b * 3 + o) * HW + pix] * (1.0f - m);
      if (out_nchw) out_nchw[b * obs + o * HW + pix] = cv;
      if (out_u8) out_u8[i * 3 + (2 - o)] = (unsigned char)(int)((cv + 1.0f) / 2.0f * 255.0f);
    }
  }
}
int head_split(const void* x, const float* w, const float* bias, int cout, int B, int H, int W, int mode, const float* img, const float* mask_bin,
               const float* mask_soft, float* out_nchw, float* out2, void* out_pack8, int no_mask_coarse, int Wp, int padl, long long obs, long long msbs,
               unsigned char* out_u8, cudaStream_t s) {
  SE_REQUIRE(cout == 1 || cout == 3, "head cout");
  const long long n = (long long)B * H * W;
  if (!obs) obs = (long long)cout * H * W;
  if (!msbs) msbs = (long long)H * W;
  if (cout == 1)
    head_split_kernel<1><<<cdiv_s(n, 128), 128, 0, s>>>((const uint4*)x, w, bias, B, H, W, mode, img, mask_bin, mask_soft, out_nchw, out2, (uint4*)out_pack8,
                                                        no_mask_coarse, Wp, padl, obs, msbs, out_u8);
  else
    head_split_kernel<3><<<cdiv_s(n, 128), 128, 0, s>>>((const uint4*)x, w, bias, B, H, W, mode, img, mask_bin, mask_soft, out_nchw, out2, (uint4*)out_pack8,
                                                        no_mask_coarse, Wp, padl, obs, msbs, out_u8);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ global pooling + broadcast
// x: [B][2*ld][HW][8] (hi blocks [0, ld), lo blocks [ld, 2 ld)); one block per (channel block, image)
__global__ void plane_reduce_split_kernel(const uint4* __restrict__ x, int ld, int C, int HW, int mode, float* __restrict__ out) {
  __shared__ float red[8][8];
  const int b = blockIdx.y, cb = blockIdx.x;
  const uint4* hi = x + ((size_t)b * 2 * ld + cb) * HW;
  const uint4* lo = hi + (size_t)ld * HW;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = (mode == RED_MAX) ? -INFINITY : 0.0f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float v[8];
    join8(hi[p], lo[p], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (mode == RED_MAX) ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    for (int o = 16; o; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, acc[i], o);
      acc[i] = (mode == RED_MAX) ? fmaxf(acc[i], t) : acc[i] + t;
    }
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < 8; ++i) red[wid][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float r = red[0][threadIdx.x];
    for (int j = 1; j < nw; ++j) r = (mode == RED_MAX) ? fmaxf(r, red[j][threadIdx.x]) : r + red[j][threadIdx.x];
    if (mode == RED_AVG) r /= (float)HW;
    const int c = cb * 8 + threadIdx.x;
    if (c < C) out[(size_t)b * C + c] = r;
  }
}
int plane_reduce_split(const void* x, int B, int HW, int C, int ld, int mode, float* out, cudaStream_t s) {
  SE_REQUIRE(mode == RED_MAX || mode == RED_AVG, "split-half reduction: max / avg");
  plane_reduce_split_kernel<<<dim3((C + 7) / 8, B), 256, 0, s>>>((const uint4*)x, ld, C, HW, mode, out);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}
// v [B][C] -> blocks [choff/8, choff/8 + C/8) of y [B][2*ld][HW][8]
__global__ void broadcast_split_kernel(const float* __restrict__ v, uint4* __restrict__ y, int C, int HW, int ld, int choff, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (b, cb, p)
  if (i >= total) return;
  const long long p = i % HW;
  long long r = i / HW;
  const int cb = (int)(r % (C >> 3));
  const long long b = r / (C >> 3);
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = v[b * C + cb * 8 + k];
  const size_t o = ((size_t)b * 2 * ld + (choff >> 3) + cb) * HW + p;
  split8(f, &y[o], &y[o + (size_t)ld * HW]);
}
int broadcast_split(const float* v, void* y, int B, int HW, int C, int ld, int choff, cudaStream_t s) {
  SE_REQUIRE(C % 8 == 0 && choff % 8 == 0, "whole channel blocks");
  const long long n = (long long)B * (C >> 3) * HW;
  broadcast_split_kernel<<<cdiv_s(n, 256), 256, 0, s>>>(v, (uint4*)y, C, HW, ld, choff, n);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ layout conversion
// NCHW fp32 -> split-half channel-blocked. layout 1: [B][2*CB][H][W][8]; layout 2 (space-to-depth): [B][4][2][CB][H/2][W/2][8];
// layout 3 (packed stem input): [B][2][H][Wp][8] with the image at [padl, padl + W). Padding channels / pixels must be zeroed first.
__global__ void nchw_to_split_kernel(const float* __restrict__ x, __half* __restrict__ y, int C, int H, int W, int layout, int Wp, int padl, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int px = (int)(i % W);
  long long r = i / W;
  const int py = (int)(r % H);
  r /= H;
  const int c = (int)(r % C);
  const long long b = r / C;
  const float v = fminf(fmaxf(x[i] * kSplitActScale, -kSplitActMax), kSplitActMax);
  const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
  const int CB = (C + 7) / 8;
  size_t oh, ol;
  if (layout == 1) {
    oh = (((size_t)b * 2 * CB + (c >> 3)) * H + py) * W + px;
    ol = oh + (size_t)CB * H * W;
  } else if (layout == 2) {
    const int Hs = H / 2, Ws = W / 2, par = (py & 1) * 2 + (px & 1);
    oh = (((size_t)b * 8 * CB + (size_t)par * 2 * CB + (c >> 3)) * Hs + (py >> 1)) * Ws + (px >> 1);
    ol = oh + (size_t)CB * Hs * Ws;
  } else {
    oh = ((size_t)b * 2 * H + py) * Wp + px + padl;
    ol = oh + (size_t)H * Wp;
  }
  y[oh * 8 + (c & 7)] = h;
  y[ol * 8 + (c & 7)] = l;
}
int nchw_to_split(const float* x, void* y, int B, int C, int H, int W, int layout, int Wp, int padl, cudaStream_t s) {
  SE_REQUIRE(layout == 1 || (layout == 2 && C % 8 == 0 && H % 2 == 0 && W % 2 == 0) || (layout == 3 && C <= 8), "split-half layout");
  const long long total = (long long)B * C * H * W;
  nchw_to_split_kernel<<<cdiv_s(total, 256), 256, 0, s>>>(x, (__half*)y, C, H, W, layout, Wp, padl, total);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}
// split-half channel-blocked [B][2*ld][HW][8] (channels [choff, choff + C)) -> fp32, NCHW (nhwc = 0) or NHWC with pixel pitch C
__global__ void split_to_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, int C, int HW, int ld, int choff, int nhwc, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long b, p;
  int c;
  if (nhwc) { c = (int)(i % C); p = (i / C) % HW; b = i / ((long long)C * HW); }
  else { p = i % HW; c = (int)((i / HW) % C); b = i / ((long long)C * HW); }
  const int cc = choff + c;
  const size_t oh = (((size_t)b * 2 * ld + (cc >> 3)) * HW + p) * 8 + (cc & 7);
  y[i] = (__half2float(x[oh]) + __half2float(x[oh + (size_t)ld * HW * 8])) * kSplitActInv;
}
int split_to_f32(const void* x, float* y, int B, int C, int HW, int ld, int choff, int nhwc, cudaStream_t s) {
  const long long total = (long long)B * C * HW;
  split_to_f32_kernel<<<cdiv_s(total, 256), 256, 0, s>>>((const __half*)x, y, C, HW, ld, choff, nhwc, total);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}
// fp32 NHWC [B][HW][C] -> split-half channel-blocked [B][2*ld][HW][8], channels [choff, choff + C)
__global__ void nhwc_f32_to_split_kernel(const float* __restrict__ x, __half* __restrict__ y, int C, int HW, int ld, int choff, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = (i / C) % HW, b = i / ((long long)C * HW);
  const float v = fminf(fmaxf(x[i] * kSplitActScale, -kSplitActMax), kSplitActMax);
  const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
  const int cc = choff + c;
  const size_t oh = (((size_t)b * 2 * ld + (cc >> 3)) * HW + p) * 8 + (cc & 7);
  y[oh] = h;
  y[oh + (size_t)ld * HW * 8] = l;
}
int nhwc_f32_to_split(const float* x, void* y, int B, int C, int HW, int ld, int choff, cudaStream_t s) {
  const long long total = (long long)B * C * HW;
  nhwc_f32_to_split_kernel<<<cdiv_s(total, 256), 256, 0, s>>>(x, (__half*)y, C, HW, ld, choff, total);
  SE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace se
