// input_proj.cu -- the input projections in front of the transformer: per backbone level a 1x1 convolution
// + GroupNorm(32, 256), and for the extra pyramid level a 3x3 / stride 2 / pad 1 convolution + GroupNorm on the last backbone
// map (/root/reference/models/memotr.py:66-78 builds them, :107-123 applies them).  Batch 1, fp32 (the reference runs them in
// fp32 with TF32 off), channel-major (C, H*W) in and out -- exactly the `srcs` layout FrameEngine.load_frame takes.
//
// A 1x1 convolution over a channel-major map is the NN GEMM  Y (C_out, P) = W (C_out, C_in) X (C_in, P) + b: both operands are
// read along their contiguous dimension, no transposition of the 1333x800-sized maps.  The 3x3 / stride-2 convolution is the
// same GEMM over an im2col buffer (C_in * 9, P_out) -- 273 output pixels at 1333x800, 20 MB.  First version: fp32 CUDA-core
// tiles (64 x 64 x 16, 4 x 4 outputs per thread); this stage is outside bench.py's step (the backbone side of the boundary).
#include "common.cuh"

namespace memotr {

__global__ void __launch_bounds__(256)
conv_gemm_nn_kernel(const float *__restrict__ W, const float *__restrict__ X, const float *__restrict__ bias, float *__restrict__ Y,
                    int M, int N, int K) {
  pdl_grid_sync();
  __shared__ float As[16][64 + 4], Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {              // A tile: 64 rows x 16 k, k contiguous in W
      const int m = i >> 4, k = i & 15;
      As[k][m] = (m0 + m < M && k0 + k < K) ? W[(long)(m0 + m) * K + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {              // B tile: 16 k x 64 pixels, pixels contiguous in X
      const int k = i >> 6, n = i & 63;
      Bs[k][n] = (k0 + k < K && n0 + n < N) ? X[(long)(k0 + k) * N + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i], b[i] = Bs[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float bv = bias ? bias[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) Y[(long)m * N + n] = acc[i][j] + bv;
    }
  }
}

// col[(c * 9 + ky * 3 + kx), oy * Wo + ox] = x[c, 2 oy + ky - 1, 2 ox + kx - 1] (0 outside): kernel 3, stride 2, padding 1
__global__ void __launch_bounds__(256)
im2col_3x3s2_kernel(const float *__restrict__ x, int C, int H, int Wd, int Ho, int Wo, float *__restrict__ col) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x, P = (long)Ho * Wo;
  if (idx >= (long)C * 9 * P) return;
  const int p = (int)(idx % P), r = (int)(idx / P), c = r / 9, t = r - c * 9, ky = t / 3, kx = t - ky * 3;
  const int oy = p / Wo, ox = p - oy * Wo, iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
  col[idx] = (iy >= 0 && iy < H && ix >= 0 && ix < Wd) ? x[((long)c * H + iy) * Wd + ix] : 0.f;
}

// GroupNorm over a channel-major (C, P) map, batch 1: the channels of a group are adjacent rows = one contiguous block of
// cpg * P floats; one CTA per group, two passes (sums in double: 134 k elements per group at level 0).  In place.
__global__ void __launch_bounds__(1024)
groupnorm_cm_kernel(float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta, int cpg, int P, float eps) {
  pdl_grid_sync();
  __shared__ double red[2][32];
  __shared__ float stat[2];
  const long n = (long)cpg * P;
  float *g = x + (long)blockIdx.x * n;
  double s = 0.0, q = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = g[i];
    s += v, q += v * v;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s, red[1][threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.0, q = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
    if (threadIdx.x == 0) {
      const double mean = s / (double)n, var = q / (double)n - mean * mean;      // biased variance, as torch.nn.GroupNorm
      stat[0] = (float)mean, stat[1] = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + (double)eps));
    }
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = blockIdx.x * cpg + (int)(i / P);
    g[i] = (g[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

}  // namespace memotr

using namespace memotr;

extern "C" int memotr_conv_gemm(const float *W, const float *X, const float *bias, float *Y, int M, int N, int K, void *stream) {
  MEMOTR_REQUIRE(W && X && Y && M > 0 && N > 0 && K > 0, "conv_gemm: bad arguments");
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  MEMOTR_LAUNCH((conv_gemm_nn_kernel), grid, 256, 0, (cudaStream_t)stream, W, X, bias, Y, M, N, K);
  return check_launch("conv_gemm");
}

extern "C" int memotr_im2col_3x3s2(const float *x, int C, int H, int Wd, float *col, void *stream) {
  MEMOTR_REQUIRE(x && col && C > 0 && H > 0 && Wd > 0, "im2col_3x3s2: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (Wd - 1) / 2 + 1;
  const long n = (long)C * 9 * Ho * Wo;
  MEMOTR_REQUIRE(n < (1L << 31) * 256, "im2col_3x3s2: too large");
  MEMOTR_LAUNCH((im2col_3x3s2_kernel), (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, x, C, H, Wd, Ho, Wo, col);
  return check_launch("im2col_3x3s2");
}

extern "C" int memotr_groupnorm_cm(float *x, const float *gamma, const float *beta, int groups, int C, int P, float eps, void *stream) {
  MEMOTR_REQUIRE(x && gamma && beta && groups > 0 && C > 0 && C % groups == 0 && P > 0, "groupnorm_cm: bad arguments");
  MEMOTR_LAUNCH((groupnorm_cm_kernel), groups, 1024, 0, (cudaStream_t)stream, x, gamma, beta, C / groups, P, eps);
  return check_launch("groupnorm_cm");
}
