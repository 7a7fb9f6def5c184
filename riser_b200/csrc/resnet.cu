// ResNet variant of the classifier (riser/nets/resnet.py).
// The reference ships no config or weights for it (train-only, hyper-parameters unpinned) and
// plausible configurations are 25k-130k parameters, so it runs as fp32 convolutions on CUDA cores:
// channel-last activations, BatchNorm folded into the weights on the host, residual add + ReLU fused
// into the convolution, per-read lengths so that ragged batches give per-read results.  The
// convolution is register-tiled (weights of a channel tile and the input rows of a position tile in
// shared memory, 8 positions x 4 channels of accumulators per thread); the per-element kernel it
// replaced is kept for shapes whose tiles do not fit shared memory.  A tcgen05 path is left for a
// later round.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <stdlib.h>

namespace riser {
namespace {

// out[b][t][co] = act( bias[co] + sum_{k,ci} w[k][ci][co] * in[b][t*stride - pad + k][ci] (+ res[b][t][co]) )
// in rows outside [0, len_in[b]) read as zero; out rows t >= len_out[b] are not written.
__global__ void __launch_bounds__(256)
conv1d_cl_kernel(const float* __restrict__ in, const int32_t* __restrict__ len_in,
                 const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ residual, float* __restrict__ out,
                 const int32_t* __restrict__ len_out, int B, int Lin_pad, int Lout_pad, int Cin, int Cout,
                 int K, int stride, int pad, int relu) {
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * Cout;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % Cout);
    const int64_t bt = idx / Cout;
    const int t = static_cast<int>(bt % Lout_pad);
    const int b = static_cast<int>(bt / Lout_pad);
    if (t >= len_out[b]) continue;
    const int Lb = len_in[b];
    const float* inb = in + static_cast<int64_t>(b) * Lin_pad * Cin;
    float acc = bias[co];
    for (int k = 0; k < K; ++k) {
      const int ti = t * stride - pad + k;
      if (ti < 0 || ti >= Lb) continue;
      const float* row = inb + static_cast<int64_t>(ti) * Cin;
      const float* wk = w + static_cast<int64_t>(k) * Cin * Cout + co;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(__ldg(row + ci), __ldg(wk + static_cast<int64_t>(ci) * Cout), acc);
    }
    if (residual) acc += residual[idx];
    out[idx] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

// The same convolution, register-tiled and persistent: a CTA keeps the weights of its channel tile ([K][Cin][CT],
// zero-padded) in shared memory for the whole launch and walks over work items (read, tile of TYP = TY * PT output
// positions), staging for each the input rows the tile touches.  Thread (tx, ty) keeps PT positions x 4 channels of
// accumulators and per (tap, input channel) does one 16-byte weight load, PT input loads (warp broadcasts) and
// 4 PT FMAs.  Geometry (TX = CT / 4, TY, the block size TX * TY rounded up to a warp) is chosen on the host so
// that the position tiles of a read waste few rows; rows outside [0, len_in) are staged as zeros and tiles beyond
// a read's output length are skipped.
template <int PT>
__global__ void __launch_bounds__(256, 3)   // 85 registers: three CTAs per SM measured faster than two (121) or four (spills)
conv1d_cl_tiled_kernel(const float* __restrict__ in, const int32_t* __restrict__ len_in,
                       const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ residual, float* __restrict__ out,
                       const int32_t* __restrict__ len_out, int B, int Lin_pad, int Lout_pad, int Cin, int Cout,
                       int K, int stride, int pad, int relu, int CT, int TX, int TY, int pitch, int tiles,
                       uint32_t cin_magic) {
  extern __shared__ __align__(16) float smem_f[];
  const int co0 = blockIdx.y * CT;
  const int TYP = TY * PT;
  const int rows = (TYP - 1) * stride + K;
  float* w_s = smem_f;                                // [K * Cin][CT]
  float* in_s = smem_f + ((K * Cin * CT + 3) & ~3);   // [rows][pitch]
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < K * Cin * CT; i += nthr) {
    const int c = i % CT, kc = i / CT;
    w_s[i] = (co0 + c < Cout) ? __ldg(w + static_cast<int64_t>(kc) * Cout + co0 + c) : 0.f;
  }
  const int tx = tid % TX, ty = tid / TX;
  const int c0 = co0 + 4 * tx;
  float bv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) bv[c] = (c0 + c < Cout) ? __ldg(bias + c0 + c) : 0.f;
  const int n_items = B * tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / tiles;
    const int t0 = (item - b * tiles) * TYP;
    const int Lo = len_out[b];
    if (t0 >= Lo) continue;                           // block-uniform: nothing of this tile is written
    const int Lb = len_in[b];
    const int ti0 = t0 * stride - pad;
    const float* inb = in + static_cast<int64_t>(b) * Lin_pad * Cin;
    __syncthreads();                                  // the previous item's rows are no longer read (and w_s is complete)
    for (int i = tid; i < rows * Cin; i += nthr) {
      const int r = (Cin == 1) ? i : static_cast<int>(__umulhi(static_cast<uint32_t>(i), cin_magic));   // i / Cin
      const int ci = i - r * Cin;
      const int ti = ti0 + r;
      in_s[r * pitch + ci] = (ti >= 0 && ti < Lb) ? __ldg(inb + static_cast<int64_t>(ti) * Cin + ci) : 0.f;
    }
    __syncthreads();
    if (ty >= TY) continue;                           // lanes that round the block up to a warp
    float acc[PT][4];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;
    const float* a_base = in_s + ty * PT * stride * pitch;
    const float* w_base = w_s + 4 * tx;
    for (int k = 0; k < K; ++k) {
      const float* a_k = a_base + k * pitch;
      const float* w_k = w_base + static_cast<size_t>(k) * Cin * CT;
#pragma unroll 4
      for (int ci = 0; ci < Cin; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(w_k + ci * CT);
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          const float a = a_k[p * stride * pitch + ci];
          acc[p][0] = fmaf(a, wv.x, acc[p][0]);
          acc[p][1] = fmaf(a, wv.y, acc[p][1]);
          acc[p][2] = fmaf(a, wv.z, acc[p][2]);
          acc[p][3] = fmaf(a, wv.w, acc[p][3]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int t = t0 + ty * PT + p;
      if (t >= Lo) break;
      const int64_t o = (static_cast<int64_t>(b) * Lout_pad + t) * Cout + c0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c0 + c >= Cout) break;
        float v = acc[p][c] + bv[c];
        if (residual) v += residual[o + c];
        out[o + c] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// MaxPool1d(kernel 2, stride 2, padding 1) (resnet.py:83): out[t] = max(in[2t-1], in[2t]), -inf padding
__global__ void __launch_bounds__(256)
maxpool_cl_kernel(const float* __restrict__ in, const int32_t* __restrict__ len_in, float* __restrict__ out,
                  const int32_t* __restrict__ len_out, int B, int Lin_pad, int Lout_pad, int C, int pad) {
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * C;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const int64_t bt = idx / C;
    const int t = static_cast<int>(bt % Lout_pad);
    const int b = static_cast<int>(bt / Lout_pad);
    if (t >= len_out[b]) continue;
    const int Lb = len_in[b];
    const float* inb = in + static_cast<int64_t>(b) * Lin_pad * C + c;
    float m = -INFINITY;
    const int t0 = 2 * t - pad;
    if (t0 >= 0 && t0 < Lb) m = inb[static_cast<int64_t>(t0) * C];
    if (t0 + 1 < Lb) m = fmaxf(m, inb[static_cast<int64_t>(t0 + 1) * C]);
    out[idx] = m;
  }
}

// The stem in one pass (resnet.py:79-84): Conv1d(1 -> C, K, stride, padding) + folded BatchNorm + ReLU followed by
// MaxPool1d(2, 2, padding 1) -- the convolution's output (the largest activation of the net) never goes to memory.
// A CTA takes kStemTile pooled positions of one read: the signal span they need is staged in shared memory
// (zeros outside [0, len_in)), a thread computes 4 channels of one pooled position = 2 conv positions x K taps.
constexpr int kStemTile = 128, kStemThreads = 256, kStemMaxK = 32;
__global__ void __launch_bounds__(kStemThreads)
stem_pool_kernel(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ len_in, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out, const int32_t* __restrict__ len_conv,
                 const int32_t* __restrict__ len_out, int B, int Lout_pad, int Cp, int K, int stride, int pad,
                 int tiles_per_read) {
  extern __shared__ __align__(16) float stem_smem[];
  float* ws = stem_smem;                        // [K][Cp]
  float* bs = ws + K * Cp;                      // [Cp]
  float* xs = bs + Cp;                          // signal span of the tile
  for (int i = threadIdx.x; i < K * Cp; i += kStemThreads) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cp; i += kStemThreads) bs[i] = bias[i];
  const int span = (2 * kStemTile - 1) * stride + K;
  const int G = Cp >> 2;
  for (int item = blockIdx.x; item < B * tiles_per_read; item += gridDim.x) {
    const int b = item / tiles_per_read;
    const int j0 = (item - b * tiles_per_read) * kStemTile;
    const int n_out = len_out[b];
    if (j0 >= n_out) continue;
    const int n_in = len_in[b], n_conv = len_conv[b];
    const int x0 = (2 * j0 - 1) * stride - pad;            // signal index of xs[0]
    __syncthreads();                                       // (weights staged / previous tile done with xs)
    const float* xb = x + static_cast<int64_t>(b) * ld_x;
    for (int i = threadIdx.x; i < span; i += kStemThreads) {
      const int t = x0 + i;
      xs[i] = (t >= 0 && t < n_in) ? __ldg(xb + t) : 0.f;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < kStemTile * G; u += kStemThreads) {
      const int jl = u / G, c = (u - jl * G) << 2;
      const int j = j0 + jl;
      if (j >= n_out) continue;
      const float4 bb = *reinterpret_cast<const float4*>(bs + c);
      float a0[4] = {bb.x, bb.y, bb.z, bb.w}, a1[4] = {bb.x, bb.y, bb.z, bb.w};
      const float* xp = xs + 2 * jl * stride;              // conv position 2j - 1 starts here, 2j `stride` further
      for (int k = 0; k < K; ++k) {
        const float4 wk = *reinterpret_cast<const float4*>(ws + k * Cp + c);
        const float v0 = xp[k], v1 = xp[k + stride];
        a0[0] = fmaf(wk.x, v0, a0[0]); a0[1] = fmaf(wk.y, v0, a0[1]); a0[2] = fmaf(wk.z, v0, a0[2]); a0[3] = fmaf(wk.w, v0, a0[3]);
        a1[0] = fmaf(wk.x, v1, a1[0]); a1[1] = fmaf(wk.y, v1, a1[1]); a1[2] = fmaf(wk.z, v1, a1[2]); a1[3] = fmaf(wk.w, v1, a1[3]);
      }
      const bool ok0 = (2 * j - 1 >= 0) && (2 * j - 1 < n_conv), ok1 = 2 * j < n_conv;     // -inf padding of the pool
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float m = -INFINITY;
        if (ok0) m = fmaxf(a0[e], 0.f);
        if (ok1) m = fmaxf(m, fmaxf(a1[e], 0.f));
        r[e] = m;
      }
      *reinterpret_cast<float4*>(out + (static_cast<int64_t>(b) * Lout_pad + j) * Cp + c) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// The same stem with the kernel size and stride known at compile time (the shapes the synthetic configurations use;
// other shapes take stem_pool_kernel): a thread computes TWO pooled positions (four conv positions) x four channels
// from a register copy of the 3 * STRIDE + K signal samples they span -- 2.4x fewer shared-memory loads per FMA and
// sixteen independent accumulators instead of eight (the generic kernel issued at 29 % of the SM's rate, waiting on
// its shared loads).  A CTA takes kStemTile2 pooled positions of one read.
constexpr int kStemTile2 = 256;
template <int K, int STRIDE>
__global__ void __launch_bounds__(kStemThreads)
stem_pool_kernel_t(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ len_in, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, const int32_t* __restrict__ len_conv,
                   const int32_t* __restrict__ len_out, int B, int Lout_pad, int Cp, int pad, int tiles_per_read) {
  extern __shared__ __align__(16) float stem_smem[];
  float* ws = stem_smem;                        // [K][Cp]
  float* bs = ws + K * Cp;                      // [Cp]
  float* xs = bs + Cp;                          // signal span of the tile
  for (int i = threadIdx.x; i < K * Cp; i += kStemThreads) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cp; i += kStemThreads) bs[i] = bias[i];
  constexpr int kSpanT = 3 * STRIDE + K;        // samples two pooled positions read
  const int span = (2 * kStemTile2 - 1) * STRIDE + K;
  const int G = Cp >> 2;
  for (int item = blockIdx.x; item < B * tiles_per_read; item += gridDim.x) {
    const int b = item / tiles_per_read;
    const int j0 = (item - b * tiles_per_read) * kStemTile2;
    const int n_out = len_out[b];
    if (j0 >= n_out) continue;
    const int n_in = len_in[b], n_conv = len_conv[b];
    const int x0 = (2 * j0 - 1) * STRIDE - pad;            // signal index of xs[0]
    __syncthreads();                                       // (weights staged / previous tile done with xs)
    const float* xb = x + static_cast<int64_t>(b) * ld_x;
    for (int i = threadIdx.x; i < span; i += kStemThreads) {
      const int t = x0 + i;
      xs[i] = (t >= 0 && t < n_in) ? __ldg(xb + t) : 0.f;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < (kStemTile2 / 2) * G; u += kStemThreads) {
      const int pl = u / G, c = (u - pl * G) << 2;
      const int j = j0 + 2 * pl;                           // pooled positions j, j + 1
      if (j >= n_out) continue;
      float xv[kSpanT];
      const float* xp = xs + 4 * pl * STRIDE;              // conv position 2j - 1 starts here
#pragma unroll
      for (int i = 0; i < kSpanT; ++i) xv[i] = xp[i];
      const float4 bb = *reinterpret_cast<const float4*>(bs + c);
      float acc[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q][0] = bb.x; acc[q][1] = bb.y; acc[q][2] = bb.z; acc[q][3] = bb.w;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 wk = *reinterpret_cast<const float4*>(ws + k * Cp + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {                      // conv positions 2j - 1 + q
          const float v = xv[k + q * STRIDE];
          acc[q][0] = fmaf(wk.x, v, acc[q][0]);
          acc[q][1] = fmaf(wk.y, v, acc[q][1]);
          acc[q][2] = fmaf(wk.z, v, acc[q][2]);
          acc[q][3] = fmaf(wk.w, v, acc[q][3]);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {                        // pooled position j + h = conv positions 2(j+h) - 1, 2(j+h)
        const int jj = j + h;
        if (jj >= n_out) break;
        const bool ok0 = (2 * jj - 1 >= 0) && (2 * jj - 1 < n_conv), ok1 = 2 * jj < n_conv;     // -inf padding of the pool
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float m = -INFINITY;
          if (ok0) m = fmaxf(acc[2 * h][e], 0.f);
          if (ok1) m = fmaxf(m, fmaxf(acc[2 * h + 1][e], 0.f));
          r[e] = m;
        }
        *reinterpret_cast<float4*>(out + (static_cast<int64_t>(b) * Lout_pad + jj) * Cp + c) = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
  }
}

// AdaptiveAvgPool1d(1) + Flatten + Linear(C, n_classes) + softmax (resnet.py:94-98, model.py:27)
__global__ void __launch_bounds__(128)
gap_linear_softmax_kernel(const float* __restrict__ in, const int32_t* __restrict__ len, const float* __restrict__ fc_w,
                          const float* __restrict__ fc_b, float* __restrict__ probs, int B, int L_pad, int C,
                          int n_classes) {
  __shared__ float feat[1024];
  __shared__ float logit[32];
  const int b = blockIdx.x;
  const int L = len[b];
  if (L <= 0) {
    for (int j = threadIdx.x; j < n_classes; j += blockDim.x) probs[b * n_classes + j] = nanf("");
    return;
  }
  const float* inb = in + static_cast<int64_t>(b) * L_pad * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < L; ++t) s += inb[static_cast<int64_t>(t) * C + c];
    feat[c] = s / static_cast<float>(L);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < n_classes; j += blockDim.x / 32) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(feat[c], fc_w[j * C + c], a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) logit[j] = a + fc_b[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = logit[0];
    for (int j = 1; j < n_classes; ++j) m = fmaxf(m, logit[j]);
    float s = 0.f;
    for (int j = 0; j < n_classes; ++j) s += expf(logit[j] - m);
    for (int j = 0; j < n_classes; ++j) probs[b * n_classes + j] = expf(logit[j] - m) / s;
  }
}

// Valid lengths after every op of the network's main chain, from the input lengths: op j is Conv1d(k, stride, pad)
// (floor((n + 2 pad - k) / stride) + 1, clamped at 0) or, k < 0, MaxPool1d(2, 2, padding pad) ((n + 2 pad) / 2: n / 2 + 1
// with the ResNet stem's padding 1, n / 2 for the ConvNet's unpadded pool; 0 stays 0).
__global__ void len_chain_kernel(const int32_t* __restrict__ len0, int B, const int32_t* __restrict__ ksp, int n_ops,
                                 int32_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = len0[b];
  for (int j = 0; j < n_ops; ++j) {
    const int k = ksp[3 * j], st = ksp[3 * j + 1], pd = ksp[3 * j + 2];
    if (k < 0) {
      n = n > 0 ? (n + 2 * pd) / 2 : 0;
    } else {
      const int num = n + 2 * pd - k;
      n = num >= 0 ? num / st + 1 : 0;
    }
    out[static_cast<int64_t>(j) * B + b] = n;
  }
}

int grid_for(int64_t total) { return static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16)); }

}  // namespace
}  // namespace riser

using namespace riser;

extern "C" int riser_conv1d_cl(const float* in, const int32_t* len_in, const float* w, const float* bias,
                               const float* residual, float* out, const int32_t* len_out, int B, int Lin_pad,
                               int Lout_pad, int Cin, int Cout, int K, int stride, int pad, int relu,
                               riser_stream_t stream) {
  RISER_REQUIRE(in && len_in && w && bias && out && len_out, "riser_conv1d_cl: null pointer");
  RISER_REQUIRE(B > 0 && Lin_pad > 0 && Lout_pad > 0 && Cin > 0 && Cout > 0 && K > 0 && stride > 0 && pad >= 0,
                "riser_conv1d_cl: bad shape");
  // register-tiled persistent kernel whenever its weight tile and input rows fit shared memory (every ResNet shape does)
  {
    constexpr int PT = 8;
    // channel tile: all output channels when <= 128 (a 67-channel layer must not pay for a second, almost empty tile)
    const int c4 = (Cout + 3) & ~3;
    const int CT = c4 <= 128 ? c4 : 64;
    const int TX = CT / 4, ty_max = 256 / TX;
    // position tiles of a read: as few as the block allows, then the smallest TY that still covers the read with them
    const int tiles = (Lout_pad + ty_max * PT - 1) / (ty_max * PT);
    const int TY = std::min(ty_max, ((Lout_pad + tiles - 1) / tiles + PT - 1) / PT);
    const int TYP = TY * PT;
    const int threads = std::min(256, (TX * TY + 31) & ~31);
    const int pitch = Cin | 1;                                    // odd row pitch: the ty's of a warp hit distinct banks
    const int rows = (TYP - 1) * stride + K;
    const size_t smem = (static_cast<size_t>((K * Cin * CT + 3) & ~3) + static_cast<size_t>(rows) * pitch) * sizeof(float);
    if (smem <= 200 * 1024 && !getenv("RISER_RESNET_NAIVE")) {
      static size_t configured[64] = {0};
      int dev = 0;
      RISER_CUDA_TRY(cudaGetDevice(&dev));
      if (smem > configured[dev & 63]) {
        RISER_CUDA_TRY(cudaFuncSetAttribute(conv1d_cl_tiled_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
        configured[dev & 63] = smem;
      }
      int per_sm = 0;
      RISER_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv1d_cl_tiled_kernel<PT>, threads, smem));
      const int64_t n_items = static_cast<int64_t>(B) * tiles;
      const int n_ct = (Cout + CT - 1) / CT;
      const int gx = static_cast<int>(std::min<int64_t>(n_items, std::max(1, 148 * std::max(per_sm, 1) / n_ct)));
      const uint32_t cin_magic = static_cast<uint32_t>((0x100000000ull + Cin - 1) / Cin);
      RISER_REQUIRE(static_cast<int64_t>(rows) * Cin < (1 << 24), "riser_conv1d_cl: tile too large");
      conv1d_cl_tiled_kernel<PT><<<dim3(gx, n_ct), threads, smem, as_stream(stream)>>>(
          in, len_in, w, bias, residual, out, len_out, B, Lin_pad, Lout_pad, Cin, Cout, K, stride, pad, relu, CT, TX, TY,
          pitch, tiles, cin_magic);
      RISER_CUDA_TRY(cudaGetLastError());
      return RISER_OK;
    }
  }
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * Cout;
  conv1d_cl_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(in, len_in, w, bias, residual, out, len_out, B,
                                                                  Lin_pad, Lout_pad, Cin, Cout, K, stride, pad, relu);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_maxpool1d_pad_cl(const float* in, const int32_t* len_in, float* out, const int32_t* len_out, int B,
                                      int Lin_pad, int Lout_pad, int C, int pad, riser_stream_t stream) {
  RISER_REQUIRE(in && len_in && out && len_out, "riser_maxpool1d_pad_cl: null pointer");
  RISER_REQUIRE(pad == 0 || pad == 1, "riser_maxpool1d_pad_cl: padding %d (0 or 1)", pad);
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * C;
  if (total == 0) return RISER_OK;
  maxpool_cl_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(in, len_in, out, len_out, B, Lin_pad, Lout_pad, C,
                                                                    pad);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_maxpool1d_cl(const float* in, const int32_t* len_in, float* out, const int32_t* len_out, int B,
                                  int Lin_pad, int Lout_pad, int C, riser_stream_t stream) {
  return riser_maxpool1d_pad_cl(in, len_in, out, len_out, B, Lin_pad, Lout_pad, C, 1, stream);
}

extern "C" int riser_stem_pool_cl(const float* x, int64_t ld_x, const int32_t* len_in, const float* w, const float* bias,
                                  float* out, const int32_t* len_conv, const int32_t* len_out, int B, int Lout_pad,
                                  int Cp, int K, int stride, int pad, riser_stream_t stream) {
  RISER_REQUIRE(x && len_in && w && bias && out && len_conv && len_out, "riser_stem_pool_cl: null pointer");
  RISER_REQUIRE(B > 0 && Lout_pad > 0 && Cp > 0 && (Cp & 3) == 0 && K > 0 && K <= kStemMaxK && stride > 0 && pad >= 0,
                "riser_stem_pool_cl: bad shape (Cp %d multiple of 4, K %d <= %d)", Cp, K, kStemMaxK);
  if (K == 19 && stride == 3) {      // compile-time shape: two pooled positions per thread from registers
    const int tiles2 = (Lout_pad + kStemTile2 - 1) / kStemTile2;
    const size_t smem2 = sizeof(float) * (static_cast<size_t>(K) * Cp + Cp + (2 * kStemTile2 - 1) * stride + K + 4);
    if (smem2 <= 48 * 1024) {
      const int grid2 = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(B) * tiles2, 148 * 8));
      stem_pool_kernel_t<19, 3><<<grid2, kStemThreads, smem2, as_stream(stream)>>>(x, ld_x, len_in, w, bias, out, len_conv,
                                                                                len_out, B, Lout_pad, Cp, pad, tiles2);
      RISER_CUDA_TRY(cudaGetLastError());
      return RISER_OK;
    }
  }
  const int tiles = (Lout_pad + kStemTile - 1) / kStemTile;
  const size_t smem = sizeof(float) * (static_cast<size_t>(K) * Cp + Cp + (2 * kStemTile - 1) * stride + K + 4);
  RISER_REQUIRE(smem <= 48 * 1024, "riser_stem_pool_cl: %zu bytes of shared memory", smem);
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(B) * tiles, 148 * 8));
  stem_pool_kernel<<<grid, kStemThreads, smem, as_stream(stream)>>>(x, ld_x, len_in, w, bias, out, len_conv, len_out, B,
                                                                     Lout_pad, Cp, K, stride, pad, tiles);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_gap_linear_softmax(const float* in, const int32_t* len, const float* fc_w, const float* fc_b,
                                        float* probs, int B, int L_pad, int C, int n_classes,
                                        riser_stream_t stream) {
  RISER_REQUIRE(in && len && fc_w && fc_b && probs, "riser_gap_linear_softmax: null pointer");
  RISER_REQUIRE(C <= 1024 && n_classes <= 32 && n_classes >= 1, "riser_gap_linear_softmax: C <= 1024, n_classes <= 32");
  gap_linear_softmax_kernel<<<B, 128, 0, as_stream(stream)>>>(in, len, fc_w, fc_b, probs, B, L_pad, C, n_classes);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_len_chain(const int32_t* len0, int B, const int32_t* ksp, int n_ops, int32_t* out,
                               riser_stream_t stream) {
  RISER_REQUIRE(B >= 0 && n_ops >= 0, "riser_len_chain: negative size");
  if (B == 0 || n_ops == 0) return RISER_OK;
  RISER_REQUIRE(len0 && ksp && out, "riser_len_chain: null pointer");
  len_chain_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(len0, B, ksp, n_ops, out);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}
