// ResNet variant of the classifier (riser/nets/resnet.py).
// The reference ships no config or weights for it (train-only, hyper-parameters unpinned) and
// plausible configurations are 25k-130k parameters, so this round it runs as generic fp32
// direct convolutions on CUDA cores: channel-last activations, BatchNorm folded into the
// weights on the host, residual add + ReLU fused into the convolution, per-read lengths so
// that ragged batches give per-read results.  A tcgen05 path is left for a later round.
#include "common.cuh"

#include <algorithm>
#include <cmath>

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

// MaxPool1d(kernel 2, stride 2, padding 1) (resnet.py:83): out[t] = max(in[2t-1], in[2t]), -inf padding
__global__ void __launch_bounds__(256)
maxpool_cl_kernel(const float* __restrict__ in, const int32_t* __restrict__ len_in, float* __restrict__ out,
                  const int32_t* __restrict__ len_out, int B, int Lin_pad, int Lout_pad, int C) {
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
    const int t0 = 2 * t - 1;
    if (t0 >= 0 && t0 < Lb) m = inb[static_cast<int64_t>(t0) * C];
    if (t0 + 1 < Lb) m = fmaxf(m, inb[static_cast<int64_t>(t0 + 1) * C]);
    out[idx] = m;
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
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * Cout;
  conv1d_cl_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(in, len_in, w, bias, residual, out, len_out, B,
                                                                  Lin_pad, Lout_pad, Cin, Cout, K, stride, pad, relu);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_maxpool1d_cl(const float* in, const int32_t* len_in, float* out, const int32_t* len_out, int B,
                                  int Lin_pad, int Lout_pad, int C, riser_stream_t stream) {
  RISER_REQUIRE(in && len_in && out && len_out, "riser_maxpool1d_cl: null pointer");
  const int64_t total = static_cast<int64_t>(B) * Lout_pad * C;
  maxpool_cl_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(in, len_in, out, len_out, B, Lin_pad, Lout_pad, C);
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
