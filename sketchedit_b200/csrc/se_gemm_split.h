// Contextual attention of the fp32-on-tensor-cores mode: split-half fp16 tcgen05 GEMMs over explicit patch matrices (se_gemm_split.cu).
#pragma once
#include "se_common.cuh"

namespace se {

struct CamSplitPlan {
  int B, h, w, C;
  int hs, ws, L;     // patch grid (4x4 patches, stride 2) and its size
  int Mp;            // L rounded up to 256: rows of every patch matrix, pitch of S
  int KQ;            // 16 * C: K of the S GEMM, N of the PV GEMM
  size_t q_bytes;    // query (= value) patches and normalised key patches, each
  size_t s_bytes, p_bytes, o_bytes;
};
int cam_split_plan(int B, int h, int w, int C, CamSplitPlan* out);

// f: fp32 NHWC [B][h][w][C]; rnorm: fp32 [B][C] (1 / plane norm); colmask: fp32 [B][L] (0 / 1 per key); out: fp32 NHWC [B][h][w][C].
// Q, Kn (q_bytes each), S (s_bytes), P (p_bytes), O (o_bytes): workspace, 128 B aligned.
int cam_forward_split(const float* f, const float* rnorm, const float* colmask, float* out, const CamSplitPlan& pl, void* Q, void* Kn, float* S, void* P,
                      float* O, cudaStream_t stream);

}  // namespace se
