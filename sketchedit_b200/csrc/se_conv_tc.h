// Host/device interface of the tcgen05 convolution (se_conv_tc.cu).
#pragma once
#include "se_common.cuh"

namespace se {

// Packed bf16 weights for the tcgen05 path:
//   row(img, tap, chunk, n) = img*img_rows + (tap*nchunks + chunk) * (n_tiles*NT) + n ,  32 K-elements per row
// i.e. a 2-D [total_rows x 32] K-major matrix that TMA slices into [NT x 32] SWIZZLE_64B slabs.
struct TcWeights {
  const void* data = nullptr;   // device, bf16
  int ntaps = 0;
  int nchunks = 0;              // 32-channel chunks per tap (Cin padded with zeros)
  int kch = 1;                  // (tap,chunk) units per pipeline stage; divides ntaps*nchunks
  int NT = 0;                   // GEMM N per tile (Cout padded to 16), <= 256
  int n_tiles = 1;
  int img_rows = 0;             // rows per image for per-image weights (attention), 0 = shared
  long long total_rows = 0;
};

struct TcParams {
  int N, Ho, Wo;
  int tiles_x, tiles_y, n_tiles;
  int stride;
  int ntaps;
  int8_t dy[MAX_TAPS], dx[MAX_TAPS];
  int nchunks, kch, NT;
  int w_rows_tc, w_img_rows;
  int num_stages;
  const float* bias;
  int Cout;
  void* y;
  int out_dt;
  int Hout, Wout, ldo, choff;
  int osy, ooy, osx, oox;
  int epi;
  float scale;
  const float* colscale;
  unsigned long long* dbg;   // optional per-CTA role timers (SE_TC_DEBUG=1)
};

int tc_choose_kch(int total_q, int NT);
int tc_plan(const ConvParams& c, const TcWeights& w, TcParams* out, int* smem_bytes);
int tc_launch(const ConvParams& c, const TcWeights& w, cudaStream_t stream);

}  // namespace se
