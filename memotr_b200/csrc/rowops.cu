// rowops.cu -- the row-wise / element-wise kernels of the hot path (everything that is not a GEMM, the sampling
// gather or the small attention).  Each replaces a chain of ATen element-wise kernels the reference issues from
// Python; the reference lines are cited per kernel.  Activations are `dtype` (f32 or bf16); geometry (reference
// points, boxes, logits, sampling locations) is always fp32.
#include "common.cuh"

namespace memotr {

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
  bf16x8_to_f32(*reinterpret_cast<const uint4 *>(p), v);
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float *p, const float (&v)[8]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[8]) {
  *reinterpret_cast<uint4 *>(p) = f32x8_to_bf16(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over C = 256 channels, one warp per row (8 channels per lane, all in registers):
//   y = LN(x [+ x2]) * gamma + beta ;  optional second output  y + pos  (the next attention's query) and an fp32 copy.
// Replaces `src = norm(src + dropout(src2))` + `with_pos_embed` (deformable_encoder.py:124-127,130 ;
// deformable_decoder.py:251-252,313-314 ; ffn.py:23-24 ; query_updater.py:126-133).  eps = 1e-5, biased variance,
// two-pass (mean, then centred sum of squares) like ATen's kernel.
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
layernorm256_kernel(const TX *__restrict__ x, int ldx, const TX *__restrict__ x2, int ldx2,
                    const float *__restrict__ gamma, const float *__restrict__ beta, float eps, TY *__restrict__ y,
                    int ldy, const TY *__restrict__ pos, int ldpos, TY *__restrict__ ypos, int ldypos,
                    float *__restrict__ y32, int ldy32, int M) {
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int c0 = (threadIdx.x & 31) * 8;
  float v[8];
  load8<TX>(x + (long)row * ldx + c0, v);
  if (x2) {
    float t[8];
    load8<TX>(x2 + (long)row * ldx2 + c0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 256.f) + eps);
  float g[8], b[8];
  load8<float>(gamma + c0, g);
  load8<float>(beta + c0, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
  store8<TY>(y + (long)row * ldy + c0, v);
  if (y32) store8<float>(y32 + (long)row * ldy32 + c0, v);
  if (ypos) {
    float p[8];
    load8<TY>(pos + (long)row * ldpos + c0, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] += v[i];  // fp32 sum, rounded once on store
    store8<TY>(ypos + (long)row * ldypos + c0, p);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Sampling locations + attention weights from the raw projections (ms_deform_attn.py:108-120):
//   ol row = [ offsets (H,L,K,2) | logits (H,L*K) ]  ->  loc (Lq,H,L,K,2), attn (Lq,H,L,K), both fp32.
//   attn = softmax over the joint L*K axis;  2-d refs: loc = ref_l + off / (W_l, H_l);  4-d refs:
//   loc = ref_l.xy + off / K * ref_l.wh * 0.5.   Reference points are rebuilt on the fly:
//   mode 0 (encoder): query q is pixel (y,x) of level lq: ref_l = ((x+.5)/(vr[lq].x W_lq), (y+.5)/(vr[lq].y H_lq)) * vr[l]
//                     (deformable_encoder.py:29-40)
//   mode 1 (decoder): ref_l = ref(q) * (vr[l].x, vr[l].y, vr[l].x, vr[l].y)   (deformable_decoder.py:82-84)
// One thread per (q, head).
__global__ void __launch_bounds__(256)
msda_prep_kernel(const float *__restrict__ ol, int ldol, const int64_t *__restrict__ shapes,
                 const int64_t *__restrict__ lsi, const float *__restrict__ vr, const float *__restrict__ ref4, int mode,
                 float *__restrict__ loc, float *__restrict__ attn, int Lq, int H, int L, int K) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Lq * H) return;
  const int q = (int)(idx / H), h = (int)(idx % H);
  const int LK = L * K;
  const float *offp = ol + (long)q * ldol + (long)h * LK * 2;
  const float *logp = ol + (long)q * ldol + (long)H * LK * 2 + (long)h * LK;

  float mx = -INFINITY;
  for (int i = 0; i < LK; ++i) mx = fmaxf(mx, logp[i]);
  float sum = 0.f;
  for (int i = 0; i < LK; ++i) sum += expf(logp[i] - mx);

  float bx = 0.f, by = 0.f, bw = 0.f, bh = 0.f;  // level-independent part of the reference point
  if (mode == 0) {
    int lq = 0;
    for (int l = 1; l < L; ++l)
      if (q >= (int)lsi[l]) lq = l;
    const int Wq = (int)shapes[2 * lq + 1], Hq = (int)shapes[2 * lq];
    const int p = q - (int)lsi[lq];
    const int y = p / Wq, x = p % Wq;
    bx = ((float)x + 0.5f) / (vr[2 * lq] * (float)Wq);
    by = ((float)y + 0.5f) / (vr[2 * lq + 1] * (float)Hq);
  } else {
    bx = ref4[4 * q], by = ref4[4 * q + 1], bw = ref4[4 * q + 2], bh = ref4[4 * q + 3];
  }
  float *locp = loc + idx * LK * 2;
  float *attp = attn + idx * LK;
  for (int l = 0; l < L; ++l) {
    const float vx = vr[2 * l], vy = vr[2 * l + 1];
    const float rx = bx * vx, ry = by * vy;
    const float Wl = (float)shapes[2 * l + 1], Hl = (float)shapes[2 * l];
    const float rw = bw * vx, rh = bh * vy;
    for (int k = 0; k < K; ++k) {
      const int i = l * K + k;
      const float ox = offp[2 * i], oy = offp[2 * i + 1];
      float lx, ly;
      if (mode == 0) {
        lx = rx + ox / Wl;
        ly = ry + oy / Hl;
      } else {
        lx = rx + ox / (float)K * rw * 0.5f;
        ly = ry + oy / (float)K * rh * 0.5f;
      }
      locp[2 * i] = lx;
      locp[2 * i + 1] = ly;
      attp[i] = expf(logp[i] - mx) / sum;
    }
  }
}


// Fast path of the same computation for L*K in {4, 8, 16, 32}: one thread per (q, head, point).  The L*K lanes of a
// (q, head) are adjacent lanes of one warp, so the softmax max / sum are xor-shuffles, and every global access is
// fully coalesced (the one-thread-per-(q,head) form above walks 128-byte rows per thread).
template <int LK>
__global__ void __launch_bounds__(256)
msda_prep_fast_kernel(const float *__restrict__ ol, int ldol, const int64_t *__restrict__ shapes,
                      const int64_t *__restrict__ lsi, const float *__restrict__ vr, const float *__restrict__ ref4,
                      int mode, float *__restrict__ loc, float *__restrict__ attn, int Lq, int H, int L, int K) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < (long)Lq * H * LK;
  const long cidx = live ? idx : (long)Lq * H * LK - 1;  // tail lanes shadow the last element (shuffles stay full-warp)
  const int i = (int)(cidx % LK);
  const long qh = cidx / LK;
  const int q = (int)(qh / H), h = (int)(qh % H);
  const float *row = ol + (long)q * ldol;
  const float2 off = *reinterpret_cast<const float2 *>(row + ((long)h * LK + i) * 2);
  const float logit = row[(long)H * LK * 2 + (long)h * LK + i];
  float mx = logit;
#pragma unroll
  for (int s = LK / 2; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  const float e = __expf(logit - mx);
  float sum = e;
#pragma unroll
  for (int s = LK / 2; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const int l = i / K;
  const float vx = __ldg(vr + 2 * l), vy = __ldg(vr + 2 * l + 1);
  // divisions: a / b is evaluated as a * rcp(b) with the correctly rounded reciprocal (<= 1 ulp from the quotient the
  // reference's torch ops produce -- far inside the 1e-4 budget; a full IEEE division costs ~10 instructions each and
  // this kernel is instruction-bound)
  float lx, ly;
  if (mode == 0) {
    int lq = 0;
    for (int t = 1; t < L; ++t)
      if (q >= (int)__ldg(lsi + t)) lq = t;
    const int Wq = (int)__ldg(shapes + 2 * lq + 1), Hq = (int)__ldg(shapes + 2 * lq);
    const int p = q - (int)__ldg(lsi + lq);
    // p / Wq without an integer division: p < 2^24 and the fractional part of (p+0.5)/Wq is >= 0.5/Wq away from an
    // integer, far above the float rounding error, so the truncation is exact
    const int y = (int)(((float)p + 0.5f) * __frcp_rn((float)Wq));
    const int x = p - y * Wq;
    const float bx = ((float)x + 0.5f) * __frcp_rn(__ldg(vr + 2 * lq) * (float)Wq);
    const float by = ((float)y + 0.5f) * __frcp_rn(__ldg(vr + 2 * lq + 1) * (float)Hq);
    lx = bx * vx + off.x * __frcp_rn((float)__ldg(shapes + 2 * l + 1));
    ly = by * vy + off.y * __frcp_rn((float)__ldg(shapes + 2 * l));
  } else {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(ref4 + 4 * q));
    const float rk = __frcp_rn((float)K);
    lx = r.x * vx + off.x * rk * (r.z * vx) * 0.5f;
    ly = r.y * vy + off.y * rk * (r.w * vy) * 0.5f;
  }
  if (live) {
    *reinterpret_cast<float2 *>(loc + idx * 2) = make_float2(lx, ly);
    attn[idx] = e * __frcp_rn(sum);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (C, HW) fp32 feature map + position map of one level -> token-major rows [row0, row0+HW) of three (S, C) buffers:
//   src_tok = src^T ;  pos_tok = pos^T + level_embed ;  q_tok = src_tok + pos_tok
// Replaces flatten(2).transpose(1,2), `pos_embed + level_embed[lvl]` and the three torch.cat of
// deformable_transformer.py:200-216 plus with_pos_embed of the first encoder layer.  32x32 tiles through smem.
template <typename T>
__global__ void __launch_bounds__(256)
tokens_kernel(const float *__restrict__ src, const float *__restrict__ pos, const float *__restrict__ lvl_embed,
              T *__restrict__ src_tok, T *__restrict__ pos_tok, T *__restrict__ q_tok, float *__restrict__ src_tok32,
              int C, int HW, int row0, int ld, const float *__restrict__ emb, const float *__restrict__ dim_i) {
  // emb != nullptr: the position map is not an input but PositionEmbeddingSine evaluated in place from the normalised
  // cumulative counts `emb` (2, HW) of pos_cumsum_kernel: channel c < C/2 from the y count, else from the x count
  pdl_grid_sync();
  __shared__ float ts[32][33], tp[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    const bool ok = c < C && p < HW;
    ts[i][tx] = ok ? src[(long)c * HW + p] : 0.f;
    tp[i][tx] = (ok && !emb) ? pos[(long)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) {
      const float s = ts[tx][i];
      float pv = tp[tx][i];
      if (emb) {
        const int npf = C >> 1, j = c >= npf ? c - npf : c;
        const float e = __fdiv_rn(emb[(long)(c >= npf) * HW + p], __ldg(dim_i + j));
        pv = (j & 1) ? cosf(e) : sinf(e);
      }
      const float pe = pv + lvl_embed[c];
      const long o = (long)(row0 + p) * ld + c;
      src_tok[o] = from_f32<T>(s);
      if (src_tok32) src_tok32[o] = s;
      pos_tok[o] = from_f32<T>(pe);
      q_tok[o] = from_f32<T>(s + pe);
    }
  }
}

// All levels in one launch, position map evaluated in place (the `emb` mode of tokens_kernel), two adjacent channels per
// thread: a sin / cos pair of PositionEmbeddingSine shares its argument (one division and one sincosf instead of two of each)
// and every store is a 4-byte bf16 pair -- 128 contiguous bytes per warp and array instead of 64.  Tile = 32 pixels x 64
// channels; blockIdx.x walks the pixel tiles of all levels (TokLevels::tile0 = prefix sums).
struct TokLevels {
  int n, hw[8], off[8], tile0[9];   // pixels, first token row and first pixel tile of every level
  const float *src[8];              // (C, H_l * W_l) fp32 feature map of every level
};

__device__ __forceinline__ void store2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16 *p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
}

template <typename T>
__global__ void __launch_bounds__(256)
tokens_levels_kernel(const TokLevels lv, const float *__restrict__ emb, const float *__restrict__ dim_i,
                     const float *__restrict__ lvl_embed, T *__restrict__ src_tok, T *__restrict__ pos_tok, T *__restrict__ q_tok,
                     float *__restrict__ src_tok32, int C, int ld) {
  pdl_grid_sync();
  __shared__ float ts[32][65];                       // [pixel][channel], odd pitch: the transposing store is conflict-free
  int l = 0;
  while (l + 1 < lv.n && (int)blockIdx.x >= lv.tile0[l + 1]) ++l;
  const int HW = lv.hw[l], p0 = ((int)blockIdx.x - lv.tile0[l]) * 32, c0 = blockIdx.y * 64;
  const float *__restrict__ src = lv.src[l];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const int p = p0 + tx;
    ts[tx][i] = p < HW ? src[(long)(c0 + i) * HW + p] : 0.f;
  }
  __syncthreads();
  const int c = c0 + 2 * tx, npf = C >> 1, half = c >= npf, j = half ? c - npf : c;     // (c even: j even, the sin channel)
  const float dj = __ldg(dim_i + j), le0 = lvl_embed[l * C + c], le1 = lvl_embed[l * C + c + 1];
  const float *__restrict__ ecnt = emb + 2 * (long)lv.off[l] + (long)half * HW;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i;
    if (p >= HW) break;
    const float s0 = ts[i][2 * tx], s1 = ts[i][2 * tx + 1];
    float sn, cs;
    sincosf(__fdiv_rn(ecnt[p], dj), &sn, &cs);
    const float pe0 = sn + le0, pe1 = cs + le1;
    const long o = (long)(lv.off[l] + p) * ld + c;
    store2(src_tok + o, s0, s1);
    if (src_tok32) store2(src_tok32 + o, s0, s1);
    store2(pos_tok + o, pe0, pe1);
    store2(q_tok + o, s0 + pe0, s1 + pe1);
  }
}

// valid ratio of one level's padding mask (deformable_transformer.py:175-190): (#valid in row 0)/W, (#valid in col 0)/H
__global__ void valid_ratio_kernel(const unsigned char *__restrict__ mask, int Hh, int Ww, float *__restrict__ out2) {
  pdl_grid_sync();
  __shared__ int cnt[2];
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  int w = 0, h = 0;
  for (int x = threadIdx.x; x < Ww; x += blockDim.x) w += mask[x] ? 0 : 1;
  for (int y = threadIdx.x; y < Hh; y += blockDim.x) h += mask[(long)y * Ww] ? 0 : 1;
  atomicAdd(&cnt[0], w);
  atomicAdd(&cnt[1], h);
  __syncthreads();
  if (threadIdx.x == 0) {
    out2[0] = (float)cnt[0] / (float)Ww;
    out2[1] = (float)cnt[1] / (float)Hh;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PositionEmbeddingSine of one level (models/position_embedding.py:23-43, normalize=True): the position map is a function
// of the padding mask alone, so it is rebuilt on the device instead of crossing PCIe with every frame (22.9 MB per
// 1333x800 frame).  Step 1: normalised cumulative counts of valid pixels down the columns (y) and along the rows (x):
//   emb = (cumsum(~mask) - 0.5) / (last + 1e-6) * scale      -> scratch (2, H, W)
__global__ void __launch_bounds__(1024)
pos_cumsum_kernel(const unsigned char *__restrict__ mask, int Hh, int Ww, float scale, float *__restrict__ emb) {
  pdl_grid_sync();
  extern __shared__ unsigned char sm_mask[];      // the level's mask: the serial counting loops below run out of shared memory
  const int HW = Hh * Ww;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) sm_mask[i] = mask[i];
  __syncthreads();
  for (int t = threadIdx.x; t < Ww + Hh; t += blockDim.x) {
    const bool col = t < Ww;                      // column t: y embedding; row t - W: x embedding
    const int n = col ? Hh : Ww, base = col ? t : (t - Ww) * Ww, step = col ? Ww : 1;
    float cnt = 0.f;
    for (int i = 0; i < n; ++i) cnt += sm_mask[base + i * step] ? 0.f : 1.f;
    const float den = cnt + 1e-6f;
    float run = 0.f;
    float *o = emb + (col ? 0 : HW);
    for (int i = 0; i < n; ++i) {
      run += sm_mask[base + i * step] ? 0.f : 1.f;
      o[base + i * step] = __fmul_rn(__fdiv_rn(run - 0.5f, den), scale);
    }
  }
}
struct PosLevels {
  int n, hw[16], off[8];   // (H, W) and first pixel of every level in the concatenated mask / scratch
};
// all levels in one launch (block = level): emb + 2 * off[l] is level l's (2, H_l * W_l) plane pair
__global__ void __launch_bounds__(1024)
pos_cumsum_levels_kernel(const unsigned char *__restrict__ mask, PosLevels lv, float scale, float *__restrict__ emb,
                         float *__restrict__ valid_ratios) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int l = blockIdx.x, Hh = lv.hw[2 * l], Ww = lv.hw[2 * l + 1], HW = Hh * Ww;
  const unsigned char *m = mask + lv.off[l];
  float *e = emb + 2L * lv.off[l];
  // the level's mask into shared memory with 16-byte loads: byte i of the level sits at sm_raw[mis + i], so the 16-byte aligned
  // interior of the global range maps onto 16-byte aligned shared-memory chunks (byte-wise loads were one dependent global
  // round trip per 1024 bytes: most of this kernel's 24 us); the unaligned head / tail bytes are copied singly
  const int mis = (int)(reinterpret_cast<uintptr_t>(m) & 15u);
  unsigned char *sm_mask = sm_raw + mis;
  const int head = mis ? min(16 - mis, HW) : 0, n16 = (HW - head) / 16, tail0 = head + n16 * 16;
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(m + head);
    uint4 *dst = reinterpret_cast<uint4 *>(sm_mask + head);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    if ((int)threadIdx.x < head) sm_mask[threadIdx.x] = m[threadIdx.x];
    if ((int)threadIdx.x < HW - tail0) sm_mask[tail0 + threadIdx.x] = m[tail0 + threadIdx.x];
  }
  __syncthreads();
  // columns (the y embedding): one thread per column walks down it -- neighbouring threads read neighbouring bytes and write
  // neighbouring floats; rows (the x embedding): one WARP per row, prefix counts by ballot + popc, 32 pixels per step with
  // coalesced stores.  (One thread per row / column with two serial passes each measured 24 us for the four levels.)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int x = threadIdx.x; x < Ww; x += blockDim.x) {
    int cnt = 0;
    for (int i = 0; i < Hh; ++i) cnt += sm_mask[i * Ww + x] ? 0 : 1;
    if (valid_ratios && x == 0) valid_ratios[2 * l + 1] = (float)cnt / (float)Hh;      // (#valid in column 0) / H
    const float den = (float)cnt + 1e-6f;
    int run = 0;
    for (int i = 0; i < Hh; ++i) {
      run += sm_mask[i * Ww + x] ? 0 : 1;
      e[i * Ww + x] = __fmul_rn(__fdiv_rn((float)run - 0.5f, den), scale);
    }
  }
  for (int y = warp; y < Hh; y += n_warps) {
    const unsigned char *row = sm_mask + y * Ww;
    int cnt = 0;
    for (int x0 = 0; x0 < Ww; x0 += 32) {
      const bool v = x0 + lane < Ww && !row[x0 + lane];
      cnt += __popc(__ballot_sync(0xffffffffu, v));
    }
    if (valid_ratios && y == 0 && lane == 0) valid_ratios[2 * l] = (float)cnt / (float)Ww;   // (#valid in row 0) / W
    const float den = (float)cnt + 1e-6f;
    int carry = 0;
    float *o = e + HW + y * Ww;
    for (int x0 = 0; x0 < Ww; x0 += 32) {
      const bool v = x0 + lane < Ww && !row[x0 + lane];
      const unsigned b = __ballot_sync(0xffffffffu, v);
      const int run = carry + __popc(b & (0xffffffffu >> (31 - lane)));
      if (x0 + lane < Ww) o[x0 + lane] = __fmul_rn(__fdiv_rn((float)run - 0.5f, den), scale);
      carry += __popc(b);
    }
  }
}
// Step 2: out[c, p] = sin / cos (emb / dim_i): channels [0, npf) from y, [npf, 2 npf) from x; even feature sin, odd cos.
// dim_i[2k] == dim_i[2k+1], so one thread produces the (sin, cos) pair of features 2k, 2k+1 of one pixel;
// blockIdx.y = (y|x half) * npf/2 + k, so no integer division is needed.
__global__ void __launch_bounds__(256)
pos_sine_kernel(const float *__restrict__ emb, const float *__restrict__ dim_i, int npf, int HW, float *__restrict__ out) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int half = blockIdx.y / (npf / 2), k = blockIdx.y % (npf / 2);
  const float e = __fdiv_rn(emb[(long)half * HW + p], __ldg(dim_i + 2 * k));
  float sv, cv;
  sincosf(e, &sv, &cv);
  float *o = out + ((long)half * npf + 2 * k) * HW + p;
  o[0] = sv;
  o[HW] = cv;
}

// ---------------------------------------------------------------------------------------------------------------
// Sine embedding of 4-d boxes (models/utils.py:78-85): out[n, c*128 + 2j] = sin(e), out[n, c*128 + 2j+1] = cos(e),
// e = p[n,c] * 2pi / dim_t[2j]  with dim_t[i] = 10000^(2*(i//2)/128) supplied by the host (computed once with the
// reference's own expression).  Optional: sigmoid first (query_updater.py:102) and a per-coordinate scale
// (the decoder's reference_points * valid_ratios, deformable_decoder.py:82-91).
template <typename T>
__global__ void __launch_bounds__(256)
sine_embed_kernel(const float *__restrict__ pts, int ldp, const float *__restrict__ scale4, int apply_sigmoid,
                  const float *__restrict__ dim_t, T *__restrict__ out, int ldo, int N) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (n, coord, j)
  if (idx >= (long)N * 256) return;
  const int n = (int)(idx >> 8), c = (int)((idx >> 6) & 3), j = (int)(idx & 63);
  float p = pts[(long)n * ldp + c];
  if (apply_sigmoid) p = 1.f / (1.f + expf(-p));
  if (scale4) p *= scale4[c];
  const float e = p * 6.283185307179586f / dim_t[2 * j];
  T *o = out + (long)n * ldo + c * 128 + 2 * j;
  o[0] = from_f32<T>(sinf(e));
  o[1] = from_f32<T>(cosf(e));
}

// out = a + b  (with_pos_embed and the q/k sums: deformable_decoder.py:246,304 ; query_updater.py:121-122)
template <typename T>
__global__ void __launch_bounds__(256)
add_kernel(const T *__restrict__ a, int lda, const T *__restrict__ b, int ldb, T *__restrict__ out, int ldo, int M,
           int N) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * N) return;
  const int r = (int)(idx / N), c = (int)(idx % N);
  out[(long)r * ldo + c] = from_f32<T>(to_f32<T>(a[(long)r * lda + c]) + to_f32<T>(b[(long)r * ldb + c]));
}

// strided copy with dtype conversion (torch.cat / .to(dtype) / slicing on the hot path)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
convert_kernel(const TS *__restrict__ src, int lds, TD *__restrict__ dst, int ldd, int M, int N) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * N) return;
  const int r = (int)(idx / N), c = (int)(idx % N);
  dst[(long)r * ldd + c] = from_f32<TD>(to_f32<TS>(src[(long)r * lds + c]));
}

__device__ __forceinline__ float inv_sigmoid(float x) {  // utils/utils.py:61-74, eps = 1e-5
  x = fminf(fmaxf(x, 0.f), 1.f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// Iterative box refinement (deformable_decoder.py:139-159): new = sigmoid(delta + inverse_sigmoid(ref));
// rows [0, n_take) of the next layer's reference take `new`, the rest keep `ref` (track queries before the merge layer).
// `new` for all rows is also the layer's predicted box (memotr.py:147-160 computes the same expression).
__global__ void box_refine_kernel(const float *__restrict__ delta, const float *__restrict__ ref,
                                  float *__restrict__ new_ref, float *__restrict__ ref_next, int N, int n_take) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 4) return;
  const float r = ref[idx];
  const float nr = sigmoidf(delta[idx] + inv_sigmoid(r));
  new_ref[idx] = nr;
  ref_next[idx] = (idx / 4 < n_take) ? nr : r;
}

// op 0: sigmoid, op 1: inverse_sigmoid  (deformable_transformer.py:241 ; memotr.py:183-187)
__global__ void unary_kernel(const float *__restrict__ in, float *__restrict__ out, long n, int op) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  out[idx] = op == 0 ? sigmoidf(in[idx]) : inv_sigmoid(in[idx]);
}

// QueryUpdater, first step (query_updater.py:84-85,99-102): is_pos = max_c sigmoid(logit) > thr;
// ref_pts[is_pos] = inverse_sigmoid(boxes[is_pos]).
__global__ void upd_prepare_kernel(const float *__restrict__ logits, int ncls, const float *__restrict__ boxes,
                                   const float *__restrict__ ref_pts, float thr, unsigned char *__restrict__ is_pos,
                                   float *__restrict__ ref_new, int Nt) {
  pdl_grid_sync();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nt) return;
  float s = -INFINITY;
  for (int c = 0; c < ncls; ++c) s = fmaxf(s, sigmoidf(logits[n * ncls + c]));
  const bool pos = s > thr;
  is_pos[n] = pos ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) ref_new[4 * n + i] = pos ? inv_sigmoid(boxes[4 * n + i]) : ref_pts[4 * n + i];
}

// QueryUpdater, last step (query_updater.py:135-147): masked state writes.
//   long_memory <- is_pos ? (1-lambda) long_memory + lambda output_embed : long_memory
//   last_output <- is_pos ? output_embed : last_output ;  query_embed <- is_pos ? query_feat : query_embed
template <typename T>
__global__ void __launch_bounds__(256)
upd_finalize_kernel(const unsigned char *__restrict__ is_pos, const T *__restrict__ feat, int ldf,
                    const float *__restrict__ out_e, float *__restrict__ query_embed, float *__restrict__ long_memory,
                    float *__restrict__ last_output, float lam, int Nt, int C) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Nt * C) return;
  const int n = (int)(idx / C), c = (int)(idx % C);
  if (!is_pos[n]) return;
  const float o = out_e[idx];
  long_memory[idx] = (1.f - lam) * long_memory[idx] + lam * o;
  last_output[idx] = o;
  query_embed[idx] = to_f32<T>(feat[(long)n * ldf + c]);
}

}  // namespace memotr

using namespace memotr;

#define MEMOTR_DTYPE_AB(dtype) MEMOTR_REQUIRE((dtype) == MEMOTR_F32 || (dtype) == MEMOTR_BF16, "dtype must be f32/bf16")

extern "C" int memotr_layernorm(const void *x, int x_dtype, int ldx, const void *x2, int ldx2, const float *gamma,
                                const float *beta, float eps, void *y, int y_dtype, int ldy, const void *pos, int ldpos,
                                void *ypos, int ldypos, float *y32, int ldy32, int M, int C, void *stream) {
  MEMOTR_REQUIRE(C == 256, "layernorm: only C == 256 is implemented (got %d)", C);
  MEMOTR_REQUIRE(M >= 0 && x && gamma && beta && y, "layernorm: bad arguments");
  MEMOTR_DTYPE_AB(x_dtype);
  MEMOTR_DTYPE_AB(y_dtype);
  MEMOTR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && (!x2 || ldx2 % 8 == 0) && (!ypos || (ldpos % 8 == 0 && ldypos % 8 == 0)) &&
                     (!y32 || ldy32 % 8 == 0) && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta),
                 "layernorm: rows must be 16-byte aligned");
  MEMOTR_REQUIRE(!ypos || pos, "layernorm: ypos requires pos");
  if (M == 0) return MEMOTR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(M, 8);
#define LN_LAUNCH(TX, TY)                                                                                         \
  MEMOTR_LAUNCH((layernorm256_kernel<TX, TY>), grid, 256, 0, st, (const TX *)x, ldx, (const TX *)x2, ldx2, gamma, beta, eps,   \
                                                    (TY *)y, ldy, (const TY *)pos, ldpos, (TY *)ypos, ldypos, y32, \
                                                    ldy32, M)
  if (x_dtype == MEMOTR_F32 && y_dtype == MEMOTR_F32) LN_LAUNCH(float, float);
  else if (x_dtype == MEMOTR_F32) LN_LAUNCH(float, __nv_bfloat16);
  else if (y_dtype == MEMOTR_BF16) LN_LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else LN_LAUNCH(__nv_bfloat16, float);
#undef LN_LAUNCH
  return check_launch("layernorm");
}

extern "C" int memotr_msda_prep(const float *ol, int ldol, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                                const float *valid_ratios, const float *ref4, int mode, float *sampling_loc,
                                float *attn_weight, int Lq, int H, int L, int K, void *stream) {
  MEMOTR_REQUIRE(Lq >= 0 && H > 0 && L > 0 && K > 0 && ol && spatial_shapes && level_start_idx && valid_ratios &&
                     sampling_loc && attn_weight,
                 "msda_prep: bad arguments");
  MEMOTR_REQUIRE(mode == 0 || (mode == 1 && ref4), "msda_prep: mode must be 0 (encoder) or 1 (decoder, needs ref4)");
  MEMOTR_REQUIRE(ldol >= 3 * H * L * K, "msda_prep: ldol too small");
  if (Lq == 0) return MEMOTR_OK;
  const int LK = L * K;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (ldol % 2 == 0) && ((reinterpret_cast<uintptr_t>(ol) & 7u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0) && (mode == 0 || aligned16(ref4));
  if (vec_ok && (LK == 4 || LK == 8 || LK == 16 || LK == 32)) {
    const long nt = (long)Lq * H * LK;
    const int grid = (int)((nt + 255) / 256);
#define PREP_FAST(N)                                                                                             \
  MEMOTR_LAUNCH((msda_prep_fast_kernel<N>), grid, 256, 0, st, ol, ldol, spatial_shapes, level_start_idx, valid_ratios, ref4, \
                                                 mode, sampling_loc, attn_weight, Lq, H, L, K)
    if (LK == 4) PREP_FAST(4);
    else if (LK == 8) PREP_FAST(8);
    else if (LK == 16) PREP_FAST(16);
    else PREP_FAST(32);
#undef PREP_FAST
    return check_launch("msda_prep_fast");
  }
  const long n = (long)Lq * H;
  MEMOTR_LAUNCH((msda_prep_kernel), (int)((n + 255) / 256), 256, 0, st, ol, ldol, spatial_shapes, level_start_idx, valid_ratios,
                                                          ref4, mode, sampling_loc, attn_weight, Lq, H, L, K);
  return check_launch("msda_prep");
}

extern "C" int memotr_tokens_from_nchw(const float *src, const float *pos, const float *level_embed, void *src_tok,
                                       void *pos_tok, void *q_tok, float *src_tok32, int C, int HW, int row0, int ld,
                                       int dtype, void *stream) {
  MEMOTR_REQUIRE(src && pos && level_embed && src_tok && pos_tok && q_tok && C > 0 && HW > 0 && ld >= C,
                 "tokens_from_nchw: bad arguments");
  MEMOTR_DTYPE_AB(dtype);
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((tokens_kernel<float>), grid, 256, 0, st, src, pos, level_embed, (float *)src_tok, (float *)pos_tok,
                                               (float *)q_tok, src_tok32, C, HW, row0, ld, (const float *)nullptr,
                                               (const float *)nullptr);
  else
    MEMOTR_LAUNCH((tokens_kernel<__nv_bfloat16>), grid, 256, 0, st, src, pos, level_embed, (__nv_bfloat16 *)src_tok,
                                                       (__nv_bfloat16 *)pos_tok, (__nv_bfloat16 *)q_tok, src_tok32, C, HW,
                                                       row0, ld, (const float *)nullptr, (const float *)nullptr);
  return check_launch("tokens_from_nchw");
}

extern "C" int memotr_tokens_from_nchw_pe(const float *src, const unsigned char *mask, int Hh, int Ww, const float *dim_i,
                                          float scale, float *scratch, const float *level_embed, void *src_tok,
                                          void *pos_tok, void *q_tok, float *src_tok32, int C, int row0, int ld, int dtype,
                                          void *stream) {
  MEMOTR_REQUIRE(src && mask && dim_i && scratch && level_embed && src_tok && pos_tok && q_tok && C > 0 && C % 4 == 0 &&
                     Hh > 0 && Ww > 0 && ld >= C,
                 "tokens_from_nchw_pe: bad arguments");
  MEMOTR_REQUIRE(Hh * Ww <= 48 * 1024, "tokens_from_nchw_pe: level larger than 48K pixels");
  MEMOTR_DTYPE_AB(dtype);
  const int HW = Hh * Ww;
  cudaStream_t st = (cudaStream_t)stream;
  MEMOTR_LAUNCH((pos_cumsum_kernel), 1, 1024, (size_t)HW, st, mask, Hh, Ww, scale, scratch);
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32));
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((tokens_kernel<float>), grid, 256, 0, st, src, (const float *)nullptr, level_embed, (float *)src_tok,
                  (float *)pos_tok, (float *)q_tok, src_tok32, C, HW, row0, ld, (const float *)scratch, dim_i);
  else
    MEMOTR_LAUNCH((tokens_kernel<__nv_bfloat16>), grid, 256, 0, st, src, (const float *)nullptr, level_embed,
                  (__nv_bfloat16 *)src_tok, (__nv_bfloat16 *)pos_tok, (__nv_bfloat16 *)q_tok, src_tok32, C, HW, row0, ld,
                  (const float *)scratch, dim_i);
  return check_launch("tokens_from_nchw_pe");
}

extern "C" int memotr_pos_cumsum_levels(const unsigned char *mask, const int *shapes_hw, const int *level_start, int n_levels,
                                        float scale, float *emb, float *valid_ratios, void *stream) {
  MEMOTR_REQUIRE(mask && shapes_hw && level_start && emb && n_levels >= 1 && n_levels <= 8, "pos_cumsum_levels: bad arguments");
  PosLevels lv;
  lv.n = n_levels;
  int mx = 0;
  for (int l = 0; l < n_levels; ++l) {
    lv.hw[2 * l] = shapes_hw[2 * l], lv.hw[2 * l + 1] = shapes_hw[2 * l + 1], lv.off[l] = level_start[l];
    MEMOTR_REQUIRE(shapes_hw[2 * l] > 0 && shapes_hw[2 * l + 1] > 0, "pos_cumsum_levels: bad level shape");
    mx = shapes_hw[2 * l] * shapes_hw[2 * l + 1] > mx ? shapes_hw[2 * l] * shapes_hw[2 * l + 1] : mx;
  }
  MEMOTR_REQUIRE(mx <= 48 * 1024, "pos_cumsum_levels: level larger than 48K pixels");
  MEMOTR_LAUNCH((pos_cumsum_levels_kernel), n_levels, 1024, (size_t)mx + 32, (cudaStream_t)stream, mask, lv, scale, emb, valid_ratios);
  return check_launch("pos_cumsum_levels");
}

// memotr_tokens_from_nchw_emb for ALL levels in one launch (C % 64 == 0, token rows 4-byte aligned pairs): `srcs` = host array
// of the levels' (C, H_l W_l) fp32 device pointers, `emb` = the planes memotr_pos_cumsum_levels wrote, level_embed (L, C).
extern "C" int memotr_tokens_from_nchw_levels(const float *const *srcs, const float *emb, const float *dim_i, const float *level_embed,
                                              void *src_tok, void *pos_tok, void *q_tok, float *src_tok32, int C,
                                              const int *shapes_hw, const int *level_start, int n_levels, int ld, int dtype,
                                              void *stream) {
  MEMOTR_REQUIRE(srcs && emb && dim_i && level_embed && src_tok && pos_tok && q_tok && shapes_hw && level_start && n_levels >= 1 &&
                     n_levels <= 8 && C > 0 && C % 64 == 0 && ld >= C && ld % 2 == 0,
                 "tokens_from_nchw_levels: bad arguments (needs C %% 64 == 0, <= 8 levels)");
  MEMOTR_DTYPE_AB(dtype);
  TokLevels lv;
  lv.n = n_levels, lv.tile0[0] = 0;
  for (int l = 0; l < n_levels; ++l) {
    MEMOTR_REQUIRE(srcs[l] && shapes_hw[2 * l] > 0 && shapes_hw[2 * l + 1] > 0, "tokens_from_nchw_levels: bad level %d", l);
    lv.hw[l] = shapes_hw[2 * l] * shapes_hw[2 * l + 1], lv.off[l] = level_start[l], lv.src[l] = srcs[l];
    lv.tile0[l + 1] = lv.tile0[l] + ceil_div(lv.hw[l], 32);
  }
  dim3 grid(lv.tile0[n_levels], C / 64);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((tokens_levels_kernel<float>), grid, 256, 0, st, lv, emb, dim_i, level_embed, (float *)src_tok, (float *)pos_tok,
                  (float *)q_tok, src_tok32, C, ld);
  else
    MEMOTR_LAUNCH((tokens_levels_kernel<__nv_bfloat16>), grid, 256, 0, st, lv, emb, dim_i, level_embed, (__nv_bfloat16 *)src_tok,
                  (__nv_bfloat16 *)pos_tok, (__nv_bfloat16 *)q_tok, src_tok32, C, ld);
  return check_launch("tokens_from_nchw_levels");
}

extern "C" int memotr_tokens_from_nchw_emb(const float *src, const float *emb, const float *dim_i, const float *level_embed,
                                           void *src_tok, void *pos_tok, void *q_tok, float *src_tok32, int C, int HW,
                                           int row0, int ld, int dtype, void *stream) {
  MEMOTR_REQUIRE(src && emb && dim_i && level_embed && src_tok && pos_tok && q_tok && C > 0 && C % 4 == 0 && HW > 0 && ld >= C,
                 "tokens_from_nchw_emb: bad arguments");
  MEMOTR_DTYPE_AB(dtype);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32));
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((tokens_kernel<float>), grid, 256, 0, st, src, (const float *)nullptr, level_embed, (float *)src_tok,
                  (float *)pos_tok, (float *)q_tok, src_tok32, C, HW, row0, ld, emb, dim_i);
  else
    MEMOTR_LAUNCH((tokens_kernel<__nv_bfloat16>), grid, 256, 0, st, src, (const float *)nullptr, level_embed,
                  (__nv_bfloat16 *)src_tok, (__nv_bfloat16 *)pos_tok, (__nv_bfloat16 *)q_tok, src_tok32, C, HW, row0, ld, emb,
                  dim_i);
  return check_launch("tokens_from_nchw_emb");
}

extern "C" int memotr_valid_ratio(const unsigned char *mask, int Hh, int Ww, float *out2, void *stream) {
  MEMOTR_REQUIRE(mask && out2 && Hh > 0 && Ww > 0, "valid_ratio: bad arguments");
  MEMOTR_LAUNCH((valid_ratio_kernel), 1, 256, 0, (cudaStream_t)stream, mask, Hh, Ww, out2);
  return check_launch("valid_ratio");
}

extern "C" int memotr_pos_embed_sine(const unsigned char *mask, int Hh, int Ww, const float *dim_i, int num_pos_feats,
                                     float scale, float *scratch, float *out, void *stream) {
  MEMOTR_REQUIRE(mask && dim_i && scratch && out && Hh > 0 && Ww > 0 && num_pos_feats > 0, "pos_embed_sine: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  MEMOTR_REQUIRE(Hh * Ww <= 48 * 1024, "pos_embed_sine: level larger than 48K pixels");
  MEMOTR_LAUNCH((pos_cumsum_kernel), 1, 1024, (size_t)Hh * Ww, st, mask, Hh, Ww, scale, scratch);
  MEMOTR_REQUIRE(num_pos_feats % 2 == 0, "pos_embed_sine: num_pos_feats must be even");
  MEMOTR_LAUNCH((pos_sine_kernel), dim3(ceil_div(Hh * Ww, 256), num_pos_feats), 256, 0, st, (const float *)scratch, dim_i,
                num_pos_feats, Hh * Ww, out);
  return check_launch("pos_embed_sine");
}

extern "C" int memotr_sine_embed(const float *pts, int ldp, const float *scale4, int apply_sigmoid, const float *dim_t,
                                 void *out, int ldo, int N, int out_dtype, void *stream) {
  MEMOTR_REQUIRE(pts && dim_t && out && N >= 0 && ldp >= 4 && ldo >= 512, "sine_embed: bad arguments");
  MEMOTR_DTYPE_AB(out_dtype);
  if (N == 0) return MEMOTR_OK;
  const long n = (long)N * 256;
  const int grid = (int)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((sine_embed_kernel<float>), grid, 256, 0, st, pts, ldp, scale4, apply_sigmoid, dim_t, (float *)out, ldo, N);
  else
    MEMOTR_LAUNCH((sine_embed_kernel<__nv_bfloat16>), grid, 256, 0, st, pts, ldp, scale4, apply_sigmoid, dim_t, (__nv_bfloat16 *)out,
                                                           ldo, N);
  return check_launch("sine_embed");
}

extern "C" int memotr_add(const void *a, int lda, const void *b, int ldb, void *out, int ldo, int M, int N, int dtype,
                          void *stream) {
  MEMOTR_REQUIRE(a && b && out && M >= 0 && N > 0, "add: bad arguments");
  MEMOTR_DTYPE_AB(dtype);
  if (M == 0) return MEMOTR_OK;
  const long n = (long)M * N;
  const int grid = (int)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((add_kernel<float>), grid, 256, 0, st, (const float *)a, lda, (const float *)b, ldb, (float *)out, ldo, M, N);
  else
    MEMOTR_LAUNCH((add_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16 *)a, lda, (const __nv_bfloat16 *)b, ldb,
                                                    (__nv_bfloat16 *)out, ldo, M, N);
  return check_launch("add");
}

extern "C" int memotr_convert(const void *src, int src_dtype, int lds, void *dst, int dst_dtype, int ldd, int M, int N,
                              void *stream) {
  MEMOTR_REQUIRE(src && dst && M >= 0 && N > 0, "convert: bad arguments");
  MEMOTR_DTYPE_AB(src_dtype);
  MEMOTR_DTYPE_AB(dst_dtype);
  if (M == 0) return MEMOTR_OK;
  const long n = (long)M * N;
  const int grid = (int)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  using bf = __nv_bfloat16;
  if (src_dtype == MEMOTR_F32 && dst_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((convert_kernel<float, float>), grid, 256, 0, st, (const float *)src, lds, (float *)dst, ldd, M, N);
  else if (src_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((convert_kernel<float, bf>), grid, 256, 0, st, (const float *)src, lds, (bf *)dst, ldd, M, N);
  else if (dst_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((convert_kernel<bf, float>), grid, 256, 0, st, (const bf *)src, lds, (float *)dst, ldd, M, N);
  else
    MEMOTR_LAUNCH((convert_kernel<bf, bf>), grid, 256, 0, st, (const bf *)src, lds, (bf *)dst, ldd, M, N);
  return check_launch("convert");
}

extern "C" int memotr_box_refine(const float *delta, const float *ref, float *new_ref, float *ref_next, int N,
                                 int n_take, void *stream) {
  MEMOTR_REQUIRE(delta && ref && new_ref && ref_next && N >= 0, "box_refine: bad arguments");
  if (N == 0) return MEMOTR_OK;
  MEMOTR_LAUNCH((box_refine_kernel), ceil_div(N * 4, 256), 256, 0, (cudaStream_t)stream, delta, ref, new_ref, ref_next, N, n_take);
  return check_launch("box_refine");
}

extern "C" int memotr_unary(const float *in, float *out, long n, int op, void *stream) {
  MEMOTR_REQUIRE(in && out && n >= 0 && (op == 0 || op == 1), "unary: bad arguments");
  if (n == 0) return MEMOTR_OK;
  MEMOTR_LAUNCH((unary_kernel), (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, in, out, n, op);
  return check_launch("unary");
}

extern "C" int memotr_upd_prepare(const float *logits, int ncls, const float *boxes, const float *ref_pts, float thr,
                                  unsigned char *is_pos, float *ref_new, int Nt, void *stream) {
  MEMOTR_REQUIRE(logits && boxes && ref_pts && is_pos && ref_new && ncls > 0 && Nt >= 0, "upd_prepare: bad arguments");
  if (Nt == 0) return MEMOTR_OK;
  MEMOTR_LAUNCH((upd_prepare_kernel), ceil_div(Nt, 128), 128, 0, (cudaStream_t)stream, logits, ncls, boxes, ref_pts, thr, is_pos,
                                                                         ref_new, Nt);
  return check_launch("upd_prepare");
}

extern "C" int memotr_upd_finalize(const unsigned char *is_pos, const void *feat, int feat_dtype, int ldf,
                                   const float *out_e, float *query_embed, float *long_memory, float *last_output,
                                   float lam, int Nt, int C, void *stream) {
  MEMOTR_REQUIRE(is_pos && feat && out_e && query_embed && long_memory && last_output && Nt >= 0 && C > 0,
                 "upd_finalize: bad arguments");
  MEMOTR_DTYPE_AB(feat_dtype);
  if (Nt == 0) return MEMOTR_OK;
  const long n = (long)Nt * C;
  const int grid = (int)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (feat_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((upd_finalize_kernel<float>), grid, 256, 0, st, is_pos, (const float *)feat, ldf, out_e, query_embed, long_memory,
                                                     last_output, lam, Nt, C);
  else
    MEMOTR_LAUNCH((upd_finalize_kernel<__nv_bfloat16>), grid, 256, 0, st, is_pos, (const __nv_bfloat16 *)feat, ldf, out_e,
                                                             query_embed, long_memory, last_output, lam, Nt, C);
  return check_launch("upd_finalize");
}
