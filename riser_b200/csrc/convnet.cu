// Network forward of the RISER hot path for sm_100a (riser/nets/cnn.py ConvNet as
// configured by riser/model/*.yaml, run by riser/model.py:22-28):
//
//   layer 0      Conv1d(1->C0,k3,same)+ReLU+MaxPool(2,2)   CUDA cores (K = 3, bandwidth bound)
//   layers 1..   Conv1d(Cin->Cout,k3,same)+ReLU+MaxPool    implicit GEMM on tcgen05 tensor
//                cores: M = flattened (read, position) rows, N = Cout, K = 3 taps x Cin.
//                Operands are fp16 tiles staged by TMA (SWIZZLE_128B), accumulators fp32 in
//                TMEM (double buffered), and bias + ReLU + max-pool + ragged-length mask are
//                a fused epilogue that writes the next layer's channel-last input.
//   head         masked global average pool / Linear(C,2) / softmax          CUDA cores
//   decide       control.py:75-82 decision codes
//
// Activation layout: layer i's input is act_i[B * Lp_i][Cp_i] fp16, channel-last, where
// read b owns rows [b*Lp_i, (b+1)*Lp_i), Lp_i even and > the longest valid length, and
// every row at or beyond a read's valid length is ZERO.  Because each read is followed by
// at least one zero row, the 'same' padding of the 3-tap convolution is obtained for free
// by loading the flat row range shifted by -1 / 0 / +1 (TMA zero-fills outside the buffer),
// and M tiles are plain 128-row slices of the flat buffer, independent of read boundaries.
#include "common.cuh"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace riser {
namespace {

constexpr int kMaxLayers = 16;
constexpr int kBlockM = 128;          // rows per tile (UMMA M)
constexpr int kBlockK = 64;           // fp16 elements per K block = 128 bytes = one swizzle row
constexpr int kMaxNTile = 256;        // UMMA N limit
constexpr int kTmemCols = 512;        // two accumulator buffers of up to 256 columns
constexpr int kConvThreads = 192;     // warp 0: TMA, warp 1: MMA, warps 2..5: epilogue
constexpr int kMinLen = 4096;         // riser/preprocess.py:8
constexpr int kDefaultConvImpl = 1;   // see riser_plan_create
constexpr int kMaxStages = 12;        // operand ring depth (deep: early layers are latency bound)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// N tile (multiple of 16, <= 256) that pads Cout the least; ties -> the larger tile.
int pick_n_tile(int cout) {
  const int c16 = round_up(cout, 16);
  if (c16 <= kMaxNTile) return c16;
  int best = 128, best_pad = round_up(cout, 128);
  for (int nt = 128; nt <= kMaxNTile; nt += 16) {
    const int p = round_up(cout, nt);
    if (p < best_pad || (p == best_pad && nt > best)) {
      best = nt;
      best_pad = p;
    }
  }
  return best;
}

struct LayerPack {
  int cin = 0, cout = 0, cin_p = 0, cout_p = 0, n_tile = 0, n_tiles = 0;
  __half* w = nullptr;    // [passes][3][cout_p][cin_p] fp16 (layers >= 1)
  float* w0 = nullptr;    // layer 0 only: fp32 [cout][3]
  float* bias = nullptr;  // fp32 [cout_p], zero padded
  float w_inv_scale = 1.f;  // weights are stored multiplied by a power of two (keeps the fp16 lo
                            // plane out of the subnormal range); the epilogue undoes it exactly
};

}  // namespace
}  // namespace riser

struct riser_model {
  int n_layers = 0;
  int precision = 0;
  int passes = 1;       // weight planes
  int act_planes = 1;   // activation planes
  int device = 0;
  int sm_count = 148;
  riser::LayerPack layer[riser::kMaxLayers];
  float* fc_w = nullptr;   // [2][c_last]
  float* fc_b = nullptr;   // [2]
  int c_last = 0;
};

namespace riser {
namespace {

struct ConvArgs {
  const float* bias;
  const int32_t* len0;
  void* out;
  int rows_in, Lp_in, Lp_out, shift;
  int cin_p, cout_p, n_tile, n_tiles, m_tiles, k_blocks, passes, stages, out_fp32;
  uint32_t idesc;
  // v2 kernel (single A load per K block, shifted descriptors per tap)
  int n_asteps;          // A loads per K block: 1, or 2 when the activation lo plane is used
  int a_plane[2];        // activation plane (0 = hi, 1 = lo) of each A step
  int n_w[2];            // weight planes multiplied against each A step
  int w_plane[2][2];
  int n_wplanes;         // weight planes resident / in the tensor
  int out_planes;        // 1, or 2 = write hi and lo fp16 planes
  int a_stages, b_stages, resident, base_off_mode;
  float w_inv_scale;
};

struct LayerPlan {
  CUtensorMap tm_a, tm_b;
  ConvArgs args;
  int grid = 0;
  size_t smem = 0;
};

}  // namespace
}  // namespace riser

struct riser_plan {
  const riser_model* model = nullptr;
  int B = 0, max_len = 0, impl = 0;
  int Lmax[riser::kMaxLayers + 1];
  int Lp[riser::kMaxLayers + 1];
  size_t act_off[riser::kMaxLayers + 1];
  char* ws = nullptr;
  riser::LayerPlan layer[riser::kMaxLayers];
};

namespace riser {
namespace {

// ------------------------------------------------------------------------------------
// layer 0: x fp32 [B, ld_x] -> act_1 [B*Lp1][cout_p] fp16.  One thread per (output row,
// 8-channel group): consecutive lanes write consecutive 16-byte chunks (coalesced).
__global__ void __launch_bounds__(256)
layer0_kernel(const float* __restrict__ x, int64_t ld_x, const int32_t* __restrict__ len0,
              const float* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ out,
              int B, int Lp1, int cout, int cout_p, int planes) {
  __shared__ float sw[64 * 3];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < cout_p * 3; i += blockDim.x) sw[i] = (i < cout * 3) ? w[i] : 0.f;
  for (int i = threadIdx.x; i < cout_p; i += blockDim.x) sb[i] = (i < cout) ? bias[i] : 0.f;
  __syncthreads();
  const uint32_t groups = cout_p >> 3;
  const uint32_t total = static_cast<uint32_t>(B) * Lp1 * groups;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t r = idx / groups;
    const int c8 = static_cast<int>(idx - r * groups);
    const int b = static_cast<int>(r / static_cast<uint32_t>(Lp1));
    const int tp = static_cast<int>(r - static_cast<uint32_t>(b) * Lp1);
    const int L = len0[b];
    uint4* o = reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * cout_p * planes) + c8;
    if (tp >= (L >> 1)) {
      *o = make_uint4(0, 0, 0, 0);
      if (planes == 2) o[groups] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const float* xr = x + static_cast<int64_t>(b) * ld_x;
    const int t = 2 * tp;
    const float xm1 = (t > 0) ? xr[t - 1] : 0.f;
    const float2 x01 = *reinterpret_cast<const float2*>(xr + t);
    const float x2 = (t + 2 < L) ? xr[t + 2] : 0.f;
    __half2 h[4], hl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c8 * 8 + 2 * j + e;
        const float w0 = sw[c * 3], w1 = sw[c * 3 + 1], w2 = sw[c * 3 + 2], bb = sb[c];
        const float y0 = fmaf(w2, x01.y, fmaf(w1, x01.x, fmaf(w0, xm1, bb)));
        const float y1 = fmaf(w2, x2, fmaf(w1, x01.y, fmaf(w0, x01.x, bb)));
        v[e] = fminf(fmaxf(fmaxf(y0, y1), 0.f), 65504.f);
      }
      h[j] = __floats2half2_rn(v[0], v[1]);
      const float2 back = __half22float2(h[j]);
      hl[j] = __floats2half2_rn(v[0] - back.x, v[1] - back.y);
    }
    *o = *reinterpret_cast<uint4*>(h);
    if (planes == 2) o[groups] = *reinterpret_cast<uint4*>(hl);
  }
}

// ------------------------------------------------------------------------------------
// tcgen05 implicit-GEMM convolution with fused bias + ReLU + MaxPool(2,2) + length mask.
struct ConvSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  float bias[2][kMaxNTile];
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

// Epilogue for one chunk of W (= 32 or 16) accumulator columns held one row per lane.
// Rows (lanes) 2j and 2j+1 are the two positions of one max-pool pair; lane parity picks
// which half of the chunk's columns this lane finishes and stores.
template <int W>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[W], const float* bias_s, int col0,
                                               bool odd, bool valid, bool writable, void* out_row,
                                               bool out_fp32, int lo_plane_off = 0, float inv_scale = 1.f) {
  constexpr int H = W / 2;
  float r[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    const float mine = __uint_as_float(odd ? v[j + H] : v[j]);
    const float send = __uint_as_float(odd ? v[j] : v[j + H]);
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    const float bsum = fmaf(fmaxf(mine, recv), inv_scale, bias_s[col0 + (odd ? H : 0) + j]);
    r[j] = valid ? fmaxf(bsum, 0.f) : 0.f;
  }
  if (!writable) return;
  const int c = col0 + (odd ? H : 0);
  if (out_fp32) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(out_row) + c);
#pragma unroll
    for (int j = 0; j < H / 4; ++j) o[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  } else {
    uint4* o = reinterpret_cast<uint4*>(static_cast<__half*>(out_row) + c);
#pragma unroll
    for (int j = 0; j < H / 8; ++j)
      o[j] = make_uint4(pack_half2(r[8 * j], r[8 * j + 1]), pack_half2(r[8 * j + 2], r[8 * j + 3]),
                        pack_half2(r[8 * j + 4], r[8 * j + 5]), pack_half2(r[8 * j + 6], r[8 * j + 7]));
    if (lo_plane_off) {   // residual plane: a = hi + lo with hi = fp16(a), lo = fp16(a - hi)
#pragma unroll
      for (int j = 0; j < H; ++j) r[j] -= __half2float(__float2half_rn(fminf(r[j], 65504.f)));
      uint4* ol = reinterpret_cast<uint4*>(static_cast<__half*>(out_row) + lo_plane_off + c);
#pragma unroll
      for (int j = 0; j < H / 8; ++j)
        ol[j] = make_uint4(pack_half2(r[8 * j], r[8 * j + 1]), pack_half2(r[8 * j + 2], r[8 * j + 3]),
                           pack_half2(r[8 * j + 4], r[8 * j + 5]), pack_half2(r[8 * j + 6], r[8 * j + 7]));
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const ConvArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  // operand ring first (1024-byte aligned for SWIZZLE_128B), bookkeeping after it
  unsigned char* base = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = kBlockM * kBlockK * 2;
  const uint32_t b_bytes = a.n_tile * kBlockK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  ConvSmem& s = *reinterpret_cast<ConvSmem*>(base + static_cast<size_t>(a.stages) * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_total = a.m_tiles * a.n_tiles;
  const int k_iters = a.passes * 3 * a.k_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&s.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int m0 = (tile / a.n_tiles) * kBlockM;
        const int n0 = (tile % a.n_tiles) * a.n_tile;
        for (int pass = 0; pass < a.passes; ++pass) {
          for (int tap = 0; tap < 3; ++tap) {
            for (int kb = 0; kb < a.k_blocks; ++kb) {
              mbar_wait(&s.empty[stage], phase ^ 1);
              unsigned char* sa = base + static_cast<size_t>(stage) * stage_bytes;
              mbar_arrive_expect_tx(&s.full[stage], stage_bytes);
              tma_load_2d(sa, &tm_a, &s.full[stage], kb * kBlockK, m0 + tap - 1);
              tma_load_2d(sa + a_bytes, &tm_b, &s.full[stage], kb * kBlockK,
                          (pass * 3 + tap) * a.cout_p + n0);
              if (++stage == a.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxNTile;
        int kb = 0;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(base + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + a_bytes);
          const int nk = min(kBlockK / 16, (a.cin_p - kb * kBlockK) / 16);
          for (int k = 0; k < nk; ++k)
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, a.idesc, (ki | k) != 0);
          umma_commit(&s.empty[stage]);            // smem slot free once these MMAs have read it
          if (ki == k_iters - 1) umma_commit(&s.tmem_full[acc]);
          if (++kb == a.k_blocks) kb = 0;
          if (++stage == a.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 64;        // 0..127
    const bool odd = lane & 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / a.n_tiles) * kBlockM;
      const int n0 = (tile % a.n_tiles) * a.n_tile;
      // stage this tile's bias slice (double buffered by acc; the barrier below orders it)
      for (int i = et; i < a.n_tile; i += 128) s.bias[acc][i] = a.bias[n0 + i];
      asm volatile("bar.sync 1, 128;" ::: "memory");

      const int r_even = (m0 + 32 * q + lane) & ~1;
      bool valid = false, writable = false;
      int64_t out_row = 0;
      if (r_even < a.rows_in) {
        const int b = r_even / a.Lp_in;
        const int tp = (r_even - b * a.Lp_in) >> 1;
        valid = tp < (a.len0[b] >> a.shift);
        writable = tp < a.Lp_out;
        out_row = static_cast<int64_t>(b) * a.Lp_out + tp;
      }
      void* orow = a.out_fp32
                       ? static_cast<void*>(static_cast<float*>(a.out) + out_row * a.cout_p + n0)
                       : static_cast<void*>(static_cast<__half*>(a.out) + out_row * a.cout_p + n0);

      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + acc * kMaxNTile;
      int col = 0;
      for (; col + 32 <= a.n_tile; col += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + col, v);
        tmem_ld_wait();
        epilogue_chunk<32>(v, s.bias[acc], col, odd, valid, writable, orow, a.out_fp32, 0, a.w_inv_scale);
      }
      if (col < a.n_tile) {
        uint32_t v[16];
        tmem_ld_32x16(t_addr + col, v);
        tmem_ld_wait();
        epilogue_chunk<16>(v, s.bias[acc], col, odd, valid, writable, orow, a.out_fp32, 0, a.w_inv_scale);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------
// v2: one A load per K block (130 rows: the 128-row tile plus one halo row on each side);
// the three taps read it through descriptors whose start address is shifted by 0/1/2 rows.
// Weights are either resident in shared memory for the whole kernel (layers whose three
// taps fit: one N tile) or streamed through their own ring.
constexpr int kARows = kBlockM + 2;
constexpr int kAStageBytes = 136 * 128;     // ring stride, multiple of 1024
constexpr int kATxBytes = kARows * 128;
constexpr int kMaxAStages = 8, kMaxBStages = 8;

struct ConvSmem2 {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  uint64_t w_full;
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  float bias[2][kMaxNTile];
};

__device__ __forceinline__ uint64_t umma_desc_sw128_off(uint32_t smem_addr, uint32_t base_offset) {
  return umma_desc_sw128(smem_addr) | (static_cast<uint64_t>(base_offset & 7) << 49);
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const ConvArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t b_bytes = a.n_tile * kBlockK * 2;
  unsigned char* a_ring = base;
  unsigned char* b_region = base + static_cast<size_t>(a.a_stages) * kAStageBytes;
  const size_t b_region_bytes = a.resident ? static_cast<size_t>(a.n_wplanes) * 3 * a.k_blocks * b_bytes
                                           : static_cast<size_t>(a.b_stages) * b_bytes;
  ConvSmem2& s = *reinterpret_cast<ConvSmem2*>(b_region + b_region_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_total = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < a.a_stages; ++i) {
      mbar_init(&s.a_full[i], 1);
      mbar_init(&s.a_empty[i], 1);
    }
    for (int i = 0; i < a.b_stages; ++i) {
      mbar_init(&s.b_full[i], 1);
      mbar_init(&s.b_empty[i], 1);
    }
    mbar_init(&s.w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tmem_full[i], 1);
      mbar_init(&s.tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&s.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (a.resident) {
        mbar_arrive_expect_tx(&s.w_full, static_cast<uint32_t>(b_region_bytes));
        for (int wp = 0; wp < a.n_wplanes; ++wp)
          for (int tap = 0; tap < 3; ++tap)
            for (int kb = 0; kb < a.k_blocks; ++kb)
              tma_load_2d(b_region + static_cast<size_t>((wp * 3 + tap) * a.k_blocks + kb) * b_bytes, &tm_b,
                          &s.w_full, kb * kBlockK, (wp * 3 + tap) * a.cout_p);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int m0 = (tile / a.n_tiles) * kBlockM;
        const int n0 = (tile % a.n_tiles) * a.n_tile;
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          for (int as = 0; as < a.n_asteps; ++as) {
            mbar_wait(&s.a_empty[sa], pa ^ 1);
            mbar_arrive_expect_tx(&s.a_full[sa], kATxBytes);
            tma_load_2d(a_ring + static_cast<size_t>(sa) * kAStageBytes, &tm_a, &s.a_full[sa],
                        a.a_plane[as] * a.cin_p + kb * kBlockK, m0 - 1);
            if (++sa == a.a_stages) {
              sa = 0;
              pa ^= 1;
            }
            if (!a.resident) {
              for (int wi = 0; wi < a.n_w[as]; ++wi)
                for (int tap = 0; tap < 3; ++tap) {
                  mbar_wait(&s.b_empty[sb], pb ^ 1);
                  mbar_arrive_expect_tx(&s.b_full[sb], b_bytes);
                  tma_load_2d(b_region + static_cast<size_t>(sb) * b_bytes, &tm_b, &s.b_full[sb],
                              kb * kBlockK, (a.w_plane[as][wi] * 3 + tap) * a.cout_p + n0);
                  if (++sb == a.b_stages) {
                    sb = 0;
                    pb ^= 1;
                  }
                }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (a.resident) {
        mbar_wait(&s.w_full, 0);
        tc_fence_after();
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxNTile;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          const int nk = min(kBlockK / 16, (a.cin_p - kb * kBlockK) / 16);
          for (int as = 0; as < a.n_asteps; ++as) {
            mbar_wait(&s.a_full[sa], pa);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_ring + static_cast<size_t>(sa) * kAStageBytes);
            for (int wi = 0; wi < a.n_w[as]; ++wi) {
              for (int tap = 0; tap < 3; ++tap) {
                uint32_t b_addr;
                if (a.resident) {
                  b_addr = smem_u32(b_region +
                                    static_cast<size_t>((a.w_plane[as][wi] * 3 + tap) * a.k_blocks + kb) * b_bytes);
                } else {
                  mbar_wait(&s.b_full[sb], pb);
                  tc_fence_after();
                  b_addr = smem_u32(b_region + static_cast<size_t>(sb) * b_bytes);
                }
                const uint64_t da = umma_desc_sw128_off(a_addr + tap * 128, a.base_off_mode ? tap : 0);
                const uint64_t db = umma_desc_sw128(b_addr);
                for (int k = 0; k < nk; ++k) {
                  umma_f16(d_tmem, da + 2 * k, db + 2 * k, a.idesc, accumulate);
                  accumulate = 1;
                }
                if (!a.resident) {
                  umma_commit(&s.b_empty[sb]);
                  if (++sb == a.b_stages) {
                    sb = 0;
                    pb ^= 1;
                  }
                }
              }
            }
            umma_commit(&s.a_empty[sa]);
            if (++sa == a.a_stages) {
              sa = 0;
              pa ^= 1;
            }
          }
        }
        umma_commit(&s.tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const bool odd = lane & 1;
    const int row_elems = a.cout_p * a.out_planes;
    const int lo_off = (a.out_planes == 2) ? a.cout_p : 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / a.n_tiles) * kBlockM;
      const int n0 = (tile % a.n_tiles) * a.n_tile;
      for (int i = et; i < a.n_tile; i += 128) s.bias[acc][i] = a.bias[n0 + i];
      asm volatile("bar.sync 1, 128;" ::: "memory");

      const int r_even = (m0 + 32 * q + lane) & ~1;
      bool valid = false, writable = false;
      int64_t out_row = 0;
      if (r_even < a.rows_in) {
        const int b = r_even / a.Lp_in;
        const int tp = (r_even - b * a.Lp_in) >> 1;
        valid = tp < (a.len0[b] >> a.shift);
        writable = tp < a.Lp_out;
        out_row = static_cast<int64_t>(b) * a.Lp_out + tp;
      }
      void* orow = a.out_fp32
                       ? static_cast<void*>(static_cast<float*>(a.out) + out_row * row_elems + n0)
                       : static_cast<void*>(static_cast<__half*>(a.out) + out_row * row_elems + n0);

      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + acc * kMaxNTile;
      int col = 0;
      for (; col + 32 <= a.n_tile; col += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + col, v);
        tmem_ld_wait();
        epilogue_chunk<32>(v, s.bias[acc], col, odd, valid, writable, orow, a.out_fp32, lo_off, a.w_inv_scale);
      }
      if (col < a.n_tile) {
        uint32_t v[16];
        tmem_ld_32x16(t_addr + col, v);
        tmem_ld_wait();
        epilogue_chunk<16>(v, s.bias[acc], col, odd, valid, writable, orow, a.out_fp32, lo_off, a.w_inv_scale);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------
// head: masked global average pool -> Linear(C, 2) -> softmax.  One CTA per read.
constexpr int kHeadThreads = 128;
__global__ void __launch_bounds__(kHeadThreads)
head_kernel(const float* __restrict__ act, const int32_t* __restrict__ len0, const float* __restrict__ fc_w,
            const float* __restrict__ fc_b, float* __restrict__ probs, float* __restrict__ feat, int B,
            int Lp, int cp, int c, int shift) {
  __shared__ float red[2][kHeadThreads / 32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = len0[b] >> shift;
  if (L <= 0) {   // shorter than 4096 samples: the reference raises in max_pool1d
    if (threadIdx.x == 0) {
      probs[2 * b] = nanf("");
      probs[2 * b + 1] = nanf("");
    }
    return;
  }
  const float* rows = act + static_cast<int64_t>(b) * Lp * cp;
  const float inv = 1.f / static_cast<float>(L);
  float l0 = 0.f, l1 = 0.f;
  for (int ch = threadIdx.x; ch < c; ch += kHeadThreads) {
    float sum = 0.f;
    for (int t = 0; t < L; ++t) sum += rows[static_cast<int64_t>(t) * cp + ch];
    const float f = sum * inv;
    if (feat) feat[static_cast<int64_t>(b) * c + ch] = f;
    l0 = fmaf(f, __ldg(fc_w + ch), l0);
    l1 = fmaf(f, __ldg(fc_w + c + ch), l1);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, d);
    l1 += __shfl_xor_sync(0xffffffffu, l1, d);
  }
  if (lane == 0) {
    red[0][warp] = l0;
    red[1][warp] = l1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    l0 = fc_b[0];
    l1 = fc_b[1];
    for (int w = 0; w < kHeadThreads / 32; ++w) {
      l0 += red[0][w];
      l1 += red[1][w];
    }
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv_s = 1.f / (e0 + e1);
    probs[2 * b] = e0 * inv_s;
    probs[2 * b + 1] = e1 * inv_s;
  }
}

// ------------------------------------------------------------------------------------
// decision rule, riser/control.py:75-82
__global__ void decide_kernel(const float* __restrict__ probs, const int32_t* __restrict__ len, int B, int M,
                              float thr, int mode, int max_len, uint8_t* __restrict__ decision) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = len[b];
  if (n <= 0) {
    decision[b] = RISER_SKIPPED;
    return;
  }
  bool any_on = false, all_off = true;
  for (int m = 0; m < M; ++m) {
    const float p_off = probs[(static_cast<int64_t>(m) * B + b) * 2];
    const float p_on = probs[(static_cast<int64_t>(m) * B + b) * 2 + 1];
    any_on |= p_on > thr;
    all_off &= p_off > thr;
  }
  uint8_t d;
  if (any_on) d = (mode == RISER_MODE_ENRICH) ? RISER_ACCEPT : RISER_REJECT;
  else if (all_off) d = (mode == RISER_MODE_DEPLETE) ? RISER_ACCEPT : RISER_REJECT;
  else if (n >= max_len) d = RISER_NO_DECISION;
  else d = RISER_TRY_AGAIN;
  decision[b] = d;
}

// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    RISER_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p)
      return fail(RISER_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return RISER_OK;
}

// fp16 [rows][cols] row-major, box (64 cols = 128 B, box_rows), SWIZZLE_128B, zero OOB fill
int make_tmap(CUtensorMap* tm, void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc;
  int st = get_encode_fn(&enc);
  if (st) return st;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(RISER_ECUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %llu box %u", static_cast<int>(r),
                static_cast<unsigned long long>(cols), static_cast<unsigned long long>(rows), box_rows);
  return RISER_OK;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void plan_lengths(const riser_model* m, int max_len, int* Lmax, int* Lp) {
  Lmax[0] = max_len;
  Lp[0] = 0;
  for (int i = 1; i <= m->n_layers; ++i) {
    Lmax[i] = Lmax[i - 1] / 2;
    Lp[i] = round_up(Lmax[i] + 1, 2);
  }
}

size_t plan_offsets(const riser_model* m, int B, const int* Lp, size_t* off) {
  size_t cur = 0;
  for (int i = 1; i <= m->n_layers; ++i) {
    off[i] = cur;
    const bool last = (i == m->n_layers);
    const size_t elt = last ? 4 : 2;
    const size_t planes = last ? 1 : m->act_planes;
    cur += align_up(static_cast<size_t>(B) * Lp[i] * m->layer[i - 1].cout_p * planes * elt, 1024);
  }
  return cur;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace
}  // namespace riser

using namespace riser;

extern "C" int riser_model_create(riser_model** out, int n_layers, const int* channels,
                                  const float* const* conv_w, const float* const* conv_b, const float* fc_w,
                                  const float* fc_b, int precision, int device) {
  RISER_REQUIRE(out && channels && conv_w && conv_b && fc_w && fc_b, "riser_model_create: null pointer");
  RISER_REQUIRE(n_layers >= 2 && n_layers <= kMaxLayers, "riser_model_create: n_layers %d outside [2, %d]",
                n_layers, kMaxLayers);
  RISER_REQUIRE(precision >= RISER_PREC_F16 && precision <= RISER_PREC_F16_X3,
                "riser_model_create: unknown precision %d", precision);
  RISER_REQUIRE(channels[0] <= 64, "riser_model_create: layer 0 supports at most 64 output channels");
  RISER_CUDA_TRY(cudaSetDevice(device));
  int major = 0;
  RISER_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(RISER_ECUDA, "device %d is sm_%dx; riser_b200 needs sm_100", device, major);

  riser_model* m = new riser_model();
  m->n_layers = n_layers;
  m->precision = precision;
  m->passes = (precision == RISER_PREC_F16) ? 1 : 2;           // weight planes (hi [, lo])
  m->act_planes = (precision == RISER_PREC_F16_X3) ? 2 : 1;    // activation planes (hi [, lo])
  m->device = device;
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
  int cin = 1, cin_p = 1;
  for (int i = 0; i < n_layers; ++i) {
    LayerPack& L = m->layer[i];
    L.cin = cin;
    L.cout = channels[i];
    L.cin_p = cin_p;
    L.n_tile = (i == 0) ? round_up(L.cout, 16) : pick_n_tile(L.cout);
    L.cout_p = round_up(L.cout, L.n_tile);
    L.n_tiles = L.cout_p / L.n_tile;
    std::vector<float> bias(L.cout_p, 0.f);
    std::memcpy(bias.data(), conv_b[i], sizeof(float) * L.cout);
    RISER_CUDA_TRY(cudaMalloc(&L.bias, sizeof(float) * L.cout_p));
    RISER_CUDA_TRY(cudaMemcpy(L.bias, bias.data(), sizeof(float) * L.cout_p, cudaMemcpyHostToDevice));
    if (i == 0) {
      RISER_CUDA_TRY(cudaMalloc(&L.w0, sizeof(float) * L.cout * 3));
      RISER_CUDA_TRY(cudaMemcpy(L.w0, conv_w[0], sizeof(float) * L.cout * 3, cudaMemcpyHostToDevice));
    } else {
      // [plane][tap][cout_p][cin_p]: tap-major so that one 2-D tensor map serves all taps
      const size_t per_pass = static_cast<size_t>(3) * L.cout_p * L.cin_p;
      std::vector<__half> w(per_pass * m->passes, __float2half(0.f));
      const float* src = conv_w[i];   // [cout][cin][3]
      float wmax = 0.f;
      for (size_t k = 0; k < static_cast<size_t>(L.cout) * L.cin * 3; ++k) wmax = std::max(wmax, std::fabs(src[k]));
      int e = 0;
      if (wmax > 0.f && std::isfinite(wmax)) {
        std::frexp(wmax, &e);          // wmax = f * 2^e, f in [0.5, 1)
        e = 12 - e;                    // scaled max magnitude in [2^11, 2^12)
        e = std::max(-24, std::min(24, e));
      }
      const float scale = std::ldexp(1.f, e);
      L.w_inv_scale = std::ldexp(1.f, -e);
      for (int co = 0; co < L.cout; ++co)
        for (int ci = 0; ci < L.cin; ++ci)
          for (int tap = 0; tap < 3; ++tap) {
            const float v = src[(static_cast<size_t>(co) * L.cin + ci) * 3 + tap] * scale;
            const __half hi = __float2half_rn(v);
            const size_t idx = (static_cast<size_t>(tap) * L.cout_p + co) * L.cin_p + ci;
            w[idx] = hi;
            if (m->passes == 2) w[per_pass + idx] = __float2half_rn(v - __half2float(hi));
          }
      RISER_CUDA_TRY(cudaMalloc(&L.w, sizeof(__half) * w.size()));
      RISER_CUDA_TRY(cudaMemcpy(L.w, w.data(), sizeof(__half) * w.size(), cudaMemcpyHostToDevice));
    }
    cin = L.cout;
    cin_p = L.cout_p;
  }
  m->c_last = channels[n_layers - 1];
  RISER_CUDA_TRY(cudaMalloc(&m->fc_w, sizeof(float) * 2 * m->c_last));
  RISER_CUDA_TRY(cudaMemcpy(m->fc_w, fc_w, sizeof(float) * 2 * m->c_last, cudaMemcpyHostToDevice));
  RISER_CUDA_TRY(cudaMalloc(&m->fc_b, sizeof(float) * 2));
  RISER_CUDA_TRY(cudaMemcpy(m->fc_b, fc_b, sizeof(float) * 2, cudaMemcpyHostToDevice));
  *out = m;
  return RISER_OK;
}

extern "C" int riser_model_destroy(riser_model* m) {
  if (!m) return RISER_OK;
  for (int i = 0; i < m->n_layers; ++i) {
    cudaFree(m->layer[i].w);
    cudaFree(m->layer[i].w0);
    cudaFree(m->layer[i].bias);
  }
  cudaFree(m->fc_w);
  cudaFree(m->fc_b);
  delete m;
  return RISER_OK;
}

extern "C" size_t riser_workspace_bytes(const riser_model* m, int B, int max_len) {
  if (!m || B <= 0 || max_len <= 0) return 0;
  int Lmax[kMaxLayers + 1], Lp[kMaxLayers + 1];
  size_t off[kMaxLayers + 1];
  plan_lengths(m, max_len, Lmax, Lp);
  return plan_offsets(m, B, Lp, off);
}

extern "C" int riser_plan_create(riser_plan** out, const riser_model* m, int B, int max_len, void* workspace,
                                 size_t workspace_bytes, riser_stream_t stream) {
  RISER_REQUIRE(out && m && workspace, "riser_plan_create: null pointer");
  RISER_REQUIRE(B > 0 && max_len >= kMinLen, "riser_plan_create: need B > 0 and max_len >= %d", kMinLen);
  RISER_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "riser_plan_create: workspace not 256-byte aligned");
  riser_plan* p = new riser_plan();
  p->model = m;
  p->B = B;
  p->max_len = max_len;
  p->ws = static_cast<char*>(workspace);
  // RISER_CONV_IMPL: 0 = v1 (one A load per tap), 1 = v2 shifted descriptors (base_offset 0),
  // 2 = v2 with base_offset = tap.  Development switch; the default is the validated one.
  p->impl = env_int("RISER_CONV_IMPL", kDefaultConvImpl);
  const bool allow_resident = env_int("RISER_CONV_RESIDENT", 1) != 0;
  plan_lengths(m, max_len, p->Lmax, p->Lp);
  const size_t need = plan_offsets(m, B, p->Lp, p->act_off);
  if (workspace_bytes < need) {
    delete p;
    return fail(RISER_ENOMEM, "riser_plan_create: workspace %zu < %zu bytes", workspace_bytes, need);
  }
  RISER_REQUIRE(static_cast<int64_t>(B) * p->Lp[1] * 8 < (int64_t(1) << 31), "riser_plan_create: B * L too large");
  RISER_REQUIRE(p->impl != 0 || m->act_planes == 1, "riser_plan_create: the v1 kernel has no X3 mode");
  RISER_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, as_stream(stream)));
  int max_smem = 0;
  RISER_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
  for (int i = 1; i < m->n_layers; ++i) {
    const LayerPack& L = m->layer[i];
    LayerPlan& lp = p->layer[i];
    const int rows_in = B * p->Lp[i];
    const bool last = (i == m->n_layers - 1);
    int st = make_tmap(&lp.tm_a, p->ws + p->act_off[i], static_cast<uint64_t>(L.cin_p) * m->act_planes, rows_in,
                       p->impl == 0 ? kBlockM : kARows);
    if (!st) st = make_tmap(&lp.tm_b, L.w, L.cin_p, static_cast<uint64_t>(m->passes) * 3 * L.cout_p, L.n_tile);
    if (st) {
      delete p;
      return st;
    }
    ConvArgs& a = lp.args;
    std::memset(&a, 0, sizeof(a));
    a.bias = L.bias;
    a.len0 = nullptr;
    a.out = p->ws + p->act_off[i + 1];
    a.rows_in = rows_in;
    a.Lp_in = p->Lp[i];
    a.Lp_out = p->Lp[i + 1];
    a.shift = i + 1;
    a.cin_p = L.cin_p;
    a.cout_p = L.cout_p;
    a.n_tile = L.n_tile;
    a.n_tiles = L.n_tiles;
    a.m_tiles = (rows_in + kBlockM - 1) / kBlockM;
    a.k_blocks = (L.cin_p + kBlockK - 1) / kBlockK;
    a.passes = m->passes;
    a.out_fp32 = last ? 1 : 0;
    a.out_planes = last ? 1 : m->act_planes;
    a.idesc = umma_idesc_f16(kBlockM, L.n_tile);
    a.w_inv_scale = L.w_inv_scale;
    const size_t b_bytes = static_cast<size_t>(L.n_tile) * kBlockK * 2;
    if (p->impl == 0) {
      const size_t stage_bytes = static_cast<size_t>(kBlockM) * kBlockK * 2 + b_bytes;
      const size_t fixed = 1024 + sizeof(ConvSmem) + 64;
      int stages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / stage_bytes);
      a.stages = std::max(2, std::min(kMaxStages, stages));
      lp.smem = fixed + a.stages * stage_bytes;
    } else {
      a.base_off_mode = (p->impl == 2) ? 1 : 0;
      a.n_wplanes = m->passes;
      // A steps: hi plane x {W_hi [, W_lo]}, then (X3) lo plane x {W_hi}
      a.n_asteps = m->act_planes;
      a.a_plane[0] = 0;
      a.n_w[0] = m->passes;
      a.w_plane[0][0] = 0;
      a.w_plane[0][1] = 1;
      a.a_plane[1] = 1;
      a.n_w[1] = 1;
      a.w_plane[1][0] = 0;
      const size_t fixed = 1024 + sizeof(ConvSmem2) + 64;
      const size_t avail = static_cast<size_t>(max_smem) - fixed;
      const size_t w_all = static_cast<size_t>(m->passes) * 3 * a.k_blocks * b_bytes;
      const int min_a = 3;
      if (allow_resident && L.n_tiles == 1 && w_all + static_cast<size_t>(min_a) * kAStageBytes <= avail) {
        a.resident = 1;
        a.b_stages = 1;
        a.a_stages = std::min<int>(kMaxAStages, static_cast<int>((avail - w_all) / kAStageBytes));
        lp.smem = fixed + w_all + static_cast<size_t>(a.a_stages) * kAStageBytes;
      } else {
        a.resident = 0;
        // split the budget: A ring gets ~1/3 (at least 2 stages), B ring the rest
        int a_st = std::max(2, std::min<int>(4, static_cast<int>(avail / 3 / kAStageBytes)));
        int b_st = static_cast<int>((avail - static_cast<size_t>(a_st) * kAStageBytes) / b_bytes);
        b_st = std::max(2, std::min(kMaxBStages, b_st));
        a_st = std::min<int>(kMaxAStages, static_cast<int>((avail - static_cast<size_t>(b_st) * b_bytes) / kAStageBytes));
        a.a_stages = a_st;
        a.b_stages = b_st;
        lp.smem = fixed + static_cast<size_t>(a_st) * kAStageBytes + static_cast<size_t>(b_st) * b_bytes;
      }
    }
    lp.grid = std::min(a.m_tiles * a.n_tiles, m->sm_count);
  }
  RISER_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  RISER_CUDA_TRY(cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  *out = p;
  return RISER_OK;
}

extern "C" int riser_plan_destroy(riser_plan* p) {
  delete p;
  return RISER_OK;
}

extern "C" int riser_forward_launches(const riser_plan* p) { return p ? p->model->n_layers + 1 : 0; }

extern "C" int riser_plan_layer_info(const riser_plan* p, int i, int64_t* offset, int* rows_per_read,
                                     int* channels_padded, int* channels, int* n_tile) {
  RISER_REQUIRE(p && offset && rows_per_read && channels_padded && channels && n_tile,
                "riser_plan_layer_info: null pointer");
  RISER_REQUIRE(i >= 1 && i <= p->model->n_layers, "riser_plan_layer_info: layer %d out of range", i);
  *offset = static_cast<int64_t>(p->act_off[i]);
  *rows_per_read = p->Lp[i];
  *channels_padded = p->model->layer[i - 1].cout_p * (i == p->model->n_layers ? 1 : p->model->act_planes);
  *channels = p->model->layer[i - 1].cout;
  *n_tile = p->model->layer[i - 1].n_tile;
  return RISER_OK;
}

extern "C" int riser_forward_stage(const riser_plan* p, int stage, const float* x, int64_t ld_x,
                                   const int32_t* len, float* probs, float* feat, riser_stream_t stream) {
  RISER_REQUIRE(p && len, "riser_forward_stage: null pointer");
  const riser_model* m = p->model;
  cudaStream_t st = as_stream(stream);
  if (stage == 0) {
    RISER_REQUIRE(x, "riser_forward_stage: null x");
    RISER_REQUIRE(ld_x >= p->max_len && (ld_x & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
                  "riser_forward: x must be 8-byte aligned with even ld_x >= max_len");
    const LayerPack& L = m->layer[0];
    const int64_t total = static_cast<int64_t>(p->B) * p->Lp[1] * (L.cout_p / 8);
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(m->sm_count) * 32));
    layer0_kernel<<<grid, 256, 0, st>>>(x, ld_x, len, L.w0, L.bias,
                                        reinterpret_cast<__half*>(p->ws + p->act_off[1]), p->B, p->Lp[1],
                                        L.cout, L.cout_p, m->act_planes);
    RISER_CUDA_TRY(cudaGetLastError());
  } else if (stage == 1) {
    for (int i = 1; i < m->n_layers; ++i) {
      const LayerPlan& lp = p->layer[i];
      ConvArgs a = lp.args;
      a.len0 = len;
      if (p->impl == 0)
        conv_tc_kernel<<<lp.grid, kConvThreads, lp.smem, st>>>(lp.tm_a, lp.tm_b, a);
      else
        conv_tc2_kernel<<<lp.grid, kConvThreads, lp.smem, st>>>(lp.tm_a, lp.tm_b, a);
      RISER_CUDA_TRY(cudaGetLastError());
    }
  } else if (stage == 2) {
    RISER_REQUIRE(probs, "riser_forward_stage: null probs");
    const int n = m->n_layers;
    head_kernel<<<p->B, kHeadThreads, 0, st>>>(reinterpret_cast<const float*>(p->ws + p->act_off[n]), len, m->fc_w,
                                               m->fc_b, probs, feat, p->B, p->Lp[n], m->layer[n - 1].cout_p,
                                               m->c_last, n);
    RISER_CUDA_TRY(cudaGetLastError());
  } else {
    return fail(RISER_EINVAL, "riser_forward_stage: stage %d outside 0..2", stage);
  }
  return RISER_OK;
}

extern "C" int riser_forward(const riser_plan* p, const float* x, int64_t ld_x, const int32_t* len, float* probs,
                             float* feat, riser_stream_t stream) {
  RISER_REQUIRE(p && x && len && probs, "riser_forward: null pointer");
  for (int stage = 0; stage < 3; ++stage) {
    const int st = riser_forward_stage(p, stage, x, ld_x, len, probs, feat, stream);
    if (st) return st;
  }
  return RISER_OK;
}

extern "C" int riser_decide(const float* probs, const int32_t* len, int B, int M, float thr, int mode,
                            int max_len, uint8_t* decision, riser_stream_t stream) {
  RISER_REQUIRE(B >= 0 && M >= 1, "riser_decide: bad B / M");
  if (B == 0) return RISER_OK;
  RISER_REQUIRE(probs && len && decision, "riser_decide: null pointer");
  RISER_REQUIRE(mode == RISER_MODE_ENRICH || mode == RISER_MODE_DEPLETE, "riser_decide: bad mode %d", mode);
  decide_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(probs, len, B, M, thr, mode, max_len, decision);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}
