// Preprocessing kernels of the RISER hot path for sm_100a:
//   riser_normalise  -- fused median / MAD / normalise / outlier smoothing over ragged
//                       int16 windows (riser/preprocess.py:108-147)
//   riser_polya_end  -- 500-sample window statistics + the 2-state poly(A) scan
//                       (riser/preprocess.py:42-79)
// Both are HBM-bound byte/integer work: one CTA per read, the window staged once in
// shared memory with 16-byte coalesced loads, exact integer selection (histogram radix
// select for whole windows, warp-shuffle bit-descent select for 500-sample windows),
// float64 arithmetic in the reference's operation order, fp32 results written with
// 16-byte stores.
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>

namespace riser {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kBins = 2048;             // first-pass histogram bins
constexpr int kBinsPerThread = kBins / kThreads;
constexpr int kSubBins = 128;           // refinement pass (shift <= 7)
constexpr int kMaxLen = 98304;          // samples staged in smem (192 KB)
constexpr int kMaxRuns = 512;           // outlier-run start indices collected per read (more: rescan path)

constexpr double kOutlierLimit = 3.5;   // riser/preprocess.py:6
constexpr double kScalingFactor = 1.4826;  // riser/preprocess.py:7

struct SelectScratch {
  uint64_t bar[2];      // one mbarrier per staging buffer (bulk-copy completion)
  uint32_t hist[kBins];
  uint32_t sub[kSubBins];
  uint32_t warp_sums[kWarps];
  uint32_t res[4];      // bin1, before1, bin2, before2
  uint32_t refined;
  uint32_t n_runs;
  int32_t red[2 * kWarps];
  int32_t runs[kMaxRuns];
};

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += warp_sums[w];
  __syncthreads();
  return base + inc - v;
}

// A window of n samples staged in shared memory at the 16-byte phase it has in global memory:
// sample i is stage[a + i], 0 <= a < 8.  for_each() hands every sample of the window to f, one
// 16-byte chunk (8 samples) per thread and step; the two edge chunks are masked.
struct StagedWindow {
  const int16_t* stage;
  int a, n;
  __device__ __forceinline__ int operator[](int i) const { return stage[a + i]; }
  template <class F>
  __device__ __forceinline__ void chunk(int c, int end, F& f) const {
    const uint4 v = reinterpret_cast<const uint4*>(stage)[c];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const int base = 8 * c;
    if (base >= a && base + 8 <= end) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f(static_cast<int>(static_cast<int16_t>(w[j] & 0xffffu)));
        f(static_cast<int>(static_cast<int16_t>(w[j] >> 16)));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = base + j;
        if (p >= a && p < end) f(static_cast<int>(static_cast<int16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffffu)));
      }
    }
  }
  // Chunk order: warp w owns the contiguous chunks [w R, (w + 1) R); lane l walks l S .. l S + S - 1 of them with
  // S odd, so the 16-byte loads of a warp stay bank-conflict free while its lanes work on samples ~8 S apart --
  // neighbouring samples of a squiggle sit on the same current level, and 32 lanes hitting the same few histogram
  // bins serialise the shared-memory atomics.  The (< 64) chunks left over per warp are taken lane by lane.
  template <class F>
  __device__ __forceinline__ void for_each(F f) const {
    const int end = a + n;
    const int n_chunks = (end + 7) >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int R = (n_chunks + kWarps - 1) / kWarps;
    const int c0 = warp * R, c1 = min(c0 + R, n_chunks);
    int S = (c1 - c0) >> 5;
    S = (S > 0) ? ((S - 1) | 1) : 0;          // largest odd number <= (c1 - c0) / 32
    for (int i = 0; i < S; ++i) chunk(c0 + lane * S + i, end, f);
    for (int c = c0 + 32 * S + lane; c < c1; c += 32) chunk(c, end, f);
  }
};

// Exact order statistics k1 <= k2 (0-based) of the keys key(sample) over the window, all <= maxkey.
// Pass A: histogram of key >> shift with shift chosen so that it fits kBins; block scan
// locates the bins holding the two ranks.  Pass B (only when shift > 0): histogram of the
// low bits inside the located bin.  Block-uniform control flow; all threads get r1, r2.
template <class KeyFn>
__device__ int select_two(KeyFn key, const StagedWindow& win, uint32_t maxkey, uint32_t k1, uint32_t k2,
                          SelectScratch& s, uint32_t& r1, uint32_t& r2, bool keep_prefix = false, int dbg = 0) {
  const int tid = threadIdx.x;
  int shift = 0;
  while ((maxkey >> shift) >= static_cast<uint32_t>(kBins)) ++shift;
  for (int i = tid; i < kBins; i += kThreads) s.hist[i] = 0;
  __syncthreads();
  if (dbg & 2) {          // timing experiment: loads and keys, no atomics
    uint32_t acc = 0;
    win.for_each([&](int e) { acc += key(e) >> shift; });
    if (acc == 0x12345678u) s.hist[0] = acc;
    if (tid == 0) s.hist[maxkey >> (shift + 1)] = win.n;
  } else {
    win.for_each([&](int e) { atomicAdd(&s.hist[key(e) >> shift], 1u); });
  }
  __syncthreads();
  uint32_t c[kBinsPerThread];
  uint32_t local = 0;
#pragma unroll
  for (int j = 0; j < kBinsPerThread; ++j) {
    c[j] = s.hist[tid * kBinsPerThread + j];
    local += c[j];
  }
  const uint32_t ex = block_exclusive_scan(local, s.warp_sums);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint32_t k = which ? k2 : k1;
    if (k >= ex && k < ex + local) {
      uint32_t run = ex;
#pragma unroll
      for (int j = 0; j < kBinsPerThread; ++j) {
        if (k >= run && k < run + c[j]) {
          s.res[2 * which] = tid * kBinsPerThread + j;
          s.res[2 * which + 1] = run;
        }
        run += c[j];
      }
    }
  }
  __syncthreads();
  const uint32_t bin1 = s.res[0], before1 = s.res[1], bin2 = s.res[2], before2 = s.res[3];
  if (shift == 0) {
    r1 = bin1;
    r2 = bin2;
    if (keep_prefix) {   // hist[j] <- number of keys <= j (used to get the MAD without a second pass)
      uint32_t run = ex;
#pragma unroll
      for (int j = 0; j < kBinsPerThread; ++j) {
        run += c[j];
        s.hist[tid * kBinsPerThread + j] = run;
      }
    }
    __syncthreads();
    return 0;
  }
  const uint32_t mask = (1u << shift) - 1u;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const uint32_t bin = which ? bin2 : bin1;
    const uint32_t kk = which ? (k2 - before2) : (k1 - before1);
    __syncthreads();
    for (int i = tid; i < kSubBins; i += kThreads) s.sub[i] = 0;
    __syncthreads();
    win.for_each([&](int e) {
      const uint32_t kx = key(e);
      if ((kx >> shift) == bin) atomicAdd(&s.sub[kx & mask], 1u);
    });
    __syncthreads();
    if (tid == 0) {
      uint32_t run = 0, found = 0;
      for (uint32_t j = 0; j <= mask; ++j) {
        if (kk >= run && kk < run + s.sub[j]) found = j;
        run += s.sub[j];
      }
      s.refined = (bin << shift) | found;
    }
    __syncthreads();
    if (which) r2 = s.refined; else r1 = s.refined;
  }
  __syncthreads();
  return shift;
}

__device__ __forceinline__ double clip_outlier(double v) {
  // riser/preprocess.py:141-147
  if (v > kOutlierLimit) return kOutlierLimit;
  if (v < -kOutlierLimit) return -kOutlierLimit;
  return v;
}

// Kernel structure (one CTA per read, persistent over reads b = blockIdx.x, + gridDim.x, ...):
//   stage    one 1-D bulk copy (cp.async.bulk, mbarrier complete_tx) of the 16-byte blocks that hold the
//            window; with two staging buffers the copy of the CTA's NEXT read is issued before the
//            current one is processed, so the global-load latency is off the critical path
//   min/max  packed 16-bit min / max over the staged chunks
//   median   histogram radix select (select_two); MAD from the same histogram by a warp-wide
//            32-ary search on the distance (or a second select when the range needed a shift)
//   normalise + smooth: every sample gets the exact quotient; the starts of outlier runs are collected
//            in shared memory and walked afterwards, one run per thread, so a rare outlier no longer
//            stalls its whole warp in a divergent slow path
__global__ void __launch_bounds__(kThreads, 5)
normalise_kernel(const int16_t* __restrict__ sig, const int64_t* __restrict__ off,
                 const int32_t* __restrict__ start, const int32_t* __restrict__ len, int B,
                 float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ med2_mad4,
                 int nbuf, int buf_samples, int max_len, int dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SelectScratch& s = *reinterpret_cast<SelectScratch*>(smem_raw);
  int16_t* stage_base = reinterpret_cast<int16_t*>(smem_raw + ((sizeof(SelectScratch) + 127) & ~size_t(127)));
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // thread 0: bulk copy of a window of nr samples at gr (whole 16-byte blocks) into staging buffer `buf`
  auto issue = [&](int nr, const int16_t* gr, int buf) {
    if (nr <= 0) return;
    const int ar = static_cast<int>((reinterpret_cast<uintptr_t>(gr) >> 1) & 7);
    const uint32_t bytes = static_cast<uint32_t>(((ar + nr + 7) >> 3) << 4);
    mbar_arrive_expect_tx(&s.bar[buf], bytes);
    bulk_load_1d(stage_base + static_cast<size_t>(buf) * buf_samples, gr - ar, bytes, &s.bar[buf]);
  };
  // one staging buffer: the NEXT read's blocks are pulled into L2 while this one is processed, so that its
  // copy at the top of the next iteration is an L2 hit and the HBM reads spread over the compute phases
  auto prefetch = [&](int nr, const int16_t* gr) {
    if (nr <= 0) return;
    const int ar = static_cast<int>((reinterpret_cast<uintptr_t>(gr) >> 1) & 7);
    bulk_prefetch_l2(gr - ar, static_cast<uint32_t>(((ar + nr + 7) >> 3) << 4));
  };
  // window length / position of the read being processed and of the CTA's next one (loaded an iteration ahead,
  // so that no global-load latency sits between the end of a read and the copy of the next)
  auto meta = [&](int r, int& nr, const int16_t*& gr) {
    nr = 0;
    gr = sig;
    if (r < B) {
      nr = min(len[r], max_len);      // memory safety: a window longer than the caller's max_len is cut to it
      gr = sig + off[r] + (start ? start[r] : 0);
    }
  };
  int n_cur, n_nxt;
  const int16_t *g_cur, *g_nxt;
  meta(blockIdx.x, n_cur, g_cur);
  if (nbuf == 2 && tid == 0) issue(n_cur, g_cur, 0);
  uint32_t phase_bits = 0;             // bit `buf` = parity the next wait on bar[buf] uses

  int it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x, ++it, n_cur = n_nxt, g_cur = g_nxt) {
    const int buf = (nbuf == 2) ? (it & 1) : 0;
    if (tid == 0 && nbuf == 1) issue(n_cur, g_cur, 0);
    meta(b + gridDim.x, n_nxt, g_nxt);
    if (tid == 0) {
      if (nbuf == 2) issue(n_nxt, g_nxt, buf ^ 1); else prefetch(n_nxt, g_nxt);
    }
    const int n = n_cur;
    if (n <= 0) continue;                 // block-uniform; nothing was issued for this read
    const int16_t* g = g_cur;
    float* o = out + static_cast<int64_t>(b) * ld_out;
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(g) >> 1) & 7);
    const int16_t* stage = stage_base + static_cast<size_t>(buf) * buf_samples;
    const StagedWindow x{stage, a, n};     // x[i] == g[i]
    mbar_wait(&s.bar[buf], (phase_bits >> buf) & 1u);
    phase_bits ^= 1u << buf;

    // ---- min / max of the window (packed 16-bit lanes for whole chunks)
    int vmin = 32767, vmax = -32768;
    {
      uint32_t mn2 = 0x7fff7fffu, mx2 = 0x80008000u;
      const uint4* s4 = reinterpret_cast<const uint4*>(stage);
      const int end = a + n;
      const int n_chunks = (end + 7) >> 3;
      for (int c = tid; c < n_chunks; c += kThreads) {
        const uint4 v = s4[c];
        const int base = 8 * c;
        if (base >= a && base + 8 <= end) {
          mn2 = __vimin3_s16x2(mn2, v.x, v.y);
          mn2 = __vimin3_s16x2(mn2, v.z, v.w);
          mx2 = __vimax3_s16x2(mx2, v.x, v.y);
          mx2 = __vimax3_s16x2(mx2, v.z, v.w);
        } else {
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int p = base + j;
            if (p >= a && p < end) {
              const int e = static_cast<int16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
              vmin = min(vmin, e);
              vmax = max(vmax, e);
            }
          }
        }
      }
      vmin = min(vmin, min(static_cast<int>(static_cast<int16_t>(mn2 & 0xffffu)), static_cast<int>(static_cast<int16_t>(mn2 >> 16))));
      vmax = max(vmax, max(static_cast<int>(static_cast<int16_t>(mx2 & 0xffffu)), static_cast<int>(static_cast<int16_t>(mx2 >> 16))));
    }
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if (lane == 0) {
      s.red[warp] = vmin;
      s.red[kWarps + warp] = vmax;
    }
    __syncthreads();
    vmin = s.red[0];
    vmax = s.red[kWarps];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) {
      vmin = min(vmin, s.red[w]);
      vmax = max(vmax, s.red[kWarps + w]);
    }

    // ---- median: mean of the two middle order statistics -> med2 = 2 * median (exact)
    const uint32_t k1 = static_cast<uint32_t>((n - 1) >> 1), k2 = static_cast<uint32_t>(n >> 1);
    uint32_t r1, r2;
    const int range = vmax - vmin;
    const int shift_used = select_two([&](int e) { return static_cast<uint32_t>(e - vmin); }, x,
                                      static_cast<uint32_t>(range), k1, k2, s, r1, r2, true, dbg);
    const int med2 = 2 * vmin + static_cast<int>(r1 + r2);

    // ---- MAD on keys d = |2x - med2| = 2|x - median| -> mad4 = 4 * MAD (exact)
    const uint32_t dmax = static_cast<uint32_t>(max(abs(2 * vmin - med2), abs(2 * vmax - med2)));
    uint32_t mad4;
    if (shift_used == 0) {
      // The value histogram already holds the whole distribution: with P[u] = #{x - vmin <= u}
      // and m = med2 - 2*vmin, #{|2x - med2| <= d} = P[floor((m+d)/2)] - P[ceil((m-d)/2) - 1];
      // the two middle order statistics of d are found by a 32-ary search on d (warps 0 and 1,
      // one probe per lane and round: 3 rounds for an 11-bit distance).
      if (warp < 2) {
        const int m = static_cast<int>(r1 + r2);
        const uint32_t want = (warp ? k2 : k1) + 1;
        int lo_d = 0, hi_d = static_cast<int>(dmax);      // invariant: count(hi_d) >= want
        while (lo_d < hi_d) {
          const int step = (hi_d - lo_d + 32) >> 5;       // ceil(span / 32)
          const int d = min(lo_d + (lane + 1) * step - 1, hi_d);
          const int hi_u = min((m + d) >> 1, range);
          const int lo_u = (m - d + 1) >> 1;              // ceil((m - d) / 2), may be <= 0
          const uint32_t cnt = s.hist[hi_u] - (lo_u > 0 ? s.hist[lo_u - 1] : 0u);
          const uint32_t okm = __ballot_sync(0xffffffffu, cnt >= want);   // lane 31 probes hi_d: never empty
          const int f = __ffs(okm) - 1;
          hi_d = min(lo_d + (f + 1) * step - 1, hi_d);
          lo_d = lo_d + f * step;
        }
        if (lane == 0) s.res[warp] = static_cast<uint32_t>(lo_d);
      }
      __syncthreads();
      mad4 = s.res[0] + s.res[1];
    } else {
      select_two([&](int e) { return static_cast<uint32_t>(abs(2 * e - med2)); }, x, dmax, k1, k2,
                 s, r1, r2);
      mad4 = r1 + r2;
    }
    if (med2_mad4 && tid == 0) {
      med2_mad4[2 * b] = med2;
      med2_mad4[2 * b + 1] = static_cast<int32_t>(mad4);
    }

    const int n_groups = (n + 3) >> 2;
    if (mad4 == 0) {   // riser/preprocess.py:123-124
      for (int gidx = tid; gidx < n_groups; gidx += kThreads) {
        const int i0 = 4 * gidx;
        if (i0 + 4 <= n) {
          *reinterpret_cast<float4*>(o + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (int i = i0; i < n; ++i) o[i] = 0.f;
        }
      }
    } else {
      const double denom = __dmul_rn(kScalingFactor, static_cast<double>(mad4) * 0.25);
      // (x - median) / denom == k / D with k = 2x - 2*median (an exact small integer) and D = 2 * denom
      // (exact).  With y = RN(1/D), q0 = RN(k*y), e = k - q0*D (exact in one FMA), RN(q0 + e*y) is the
      // correctly rounded quotient (Markstein's final division step), i.e. bit-identical to the reference's
      // float64 divide at 3 FP64 ops instead of a full DDIV.
      const double Dd = __dmul_rn(2.0, denom);
      const double yrcp = __ddiv_rn(1.0, Dd);
      auto quot = [&](int ki) {
        const double k = static_cast<double>(ki);
        const double q0 = __dmul_rn(k, yrcp);
        const double e = __fma_rn(-q0, Dd, k);
        return __fma_rn(e, yrcp, q0);
      };
      // A window holds at most range + 1 distinct values (~1,000 for a squiggle against 4,096-16,000 samples),
      // and the float64 pipe is what bounds this kernel: when the value range fits the histogram, the fp32 result
      // of every VALUE is tabulated once (in the histogram's storage, free after the MAD search) and the per-sample
      // work becomes one shared-memory look-up.
      const bool use_lut = (shift_used == 0) && !(dbg & 8);
      float* lut = reinterpret_cast<float*>(s.hist);
      if (use_lut)
        for (int u = tid; u <= range; u += kThreads) lut[u] = static_cast<float>(quot(2 * (vmin + u) - med2));

      // ---- outlier threshold in the integer key domain: |z| > 3.5  <=>  d > dthr, where
      //      z = fl((d/2) / denom) is monotone in d.  Found once per read with exact divides: the lanes of
      //      warp 0 test the 32 candidates around 7 * denom at once (serial search as a fallback).
      if (warp == 0) {
        auto zval = [&](int d) { return __ddiv_rn(static_cast<double>(d) * 0.5, denom); };
        const int base = static_cast<int>(__dmul_rn(7.0, denom));
        const int d = base - 8 + lane;
        const uint32_t okm = __ballot_sync(0xffffffffu, d < 0 || zval(d) <= kOutlierLimit);
        if (lane == 0) {
          const int cnt = __popc(okm);
          int d0;
          if (cnt > 0 && cnt < 32 && okm == ((1u << cnt) - 1u)) {
            d0 = base - 9 + cnt;
          } else {
            d0 = base;
            while (zval(d0 + 1) <= kOutlierLimit) ++d0;
            while (d0 > 0 && zval(d0) > kOutlierLimit) --d0;
          }
          s.refined = static_cast<uint32_t>(max(d0, 0));
          s.n_runs = 0;
        }
      }
      __syncthreads();
      const uint32_t dthr = s.refined;
      auto flagged = [&](int i) { return static_cast<uint32_t>(abs(2 * x[i] - med2)) > dthr; };
      // riser/preprocess.py:130-138 for the run of consecutive outliers that starts at i: arr[j-1] is
      // already updated, arr[j+1] is still raw; the end points of the array are not clipped.
      // The recurrence is sequential, and a float64 op has a long dependent latency: per step only ONE fma
      // (prev / 2 + raw / 2, the same rounding as (prev + raw) / 2) and the clip stay on the chain; the raw
      // neighbour's quotient and the "is the next sample an outlier" test are computed a step ahead.
      auto walk = [&](int i) {
        double prev = (i > 0) ? quot(2 * x[i - 1] - med2) : 0.0;
        double rh_next = 0.5 * quot(2 * x[min(i + 1, n - 1)] - med2);        // raw value of sample j + 1, halved
        for (int j = i;; ++j) {
          const double rh = rh_next;
          const bool more = (j + 1 < n) && flagged(j + 1);
          if (more) rh_next = 0.5 * quot(2 * x[min(j + 2, n - 1)] - med2);
          double nv;
          if (j == 0) {
            nv = 2.0 * rh;                       // arr[0] = arr[1] (raw)
          } else if (j == n - 1) {
            nv = prev;                           // arr[n-1] = arr[n-2] (already updated)
          } else {
            nv = clip_outlier(__fma_rn(prev, 0.5, rh));
          }
          o[j] = static_cast<float>(nv);
          prev = nv;
          if (!more) break;
        }
      };

      // ---- normalise: 4 samples per thread and step (two 8-byte shared loads realigned by the window's
      //      phase), 16-byte stores; outlier-run starts are pushed to s.runs
      const int sh = a & 3;
      const int16_t* q8 = stage + (a & ~3);
      // |2x - med2| > dthr  <=>  x > x_hi or x < x_lo (clamped to int16: beyond the type nothing can exceed);
      // tested for the group's four samples at once on the packed 16-bit halves
      const int x_hi = min(32767, (med2 + static_cast<int>(dthr)) >> 1);            // floor
      const int x_lo = max(-32768, (med2 - static_cast<int>(dthr) + 1) >> 1);       // ceil
      const uint32_t hi2 = (static_cast<uint32_t>(x_hi) & 0xffffu) * 0x10001u;
      const uint32_t lo2 = (static_cast<uint32_t>(x_lo) & 0xffffu) * 0x10001u;
      // Outlier bookkeeping stays out of the hot loop: a thread only notes WHICH of its groups held an outlier (one
      // bit per step) and looks at those groups again afterwards, pushing the starts of outlier runs to s.runs.
      auto note_runs = [&](uint32_t mask, int step0) {
        while (mask) {
          const int it_ = __ffs(mask) - 1;
          mask &= mask - 1;
          const int i0 = 4 * (tid + (step0 + it_) * kThreads);
          const int cnt = min(4, n - i0);
          bool fprev = (i0 > 0) && flagged(i0 - 1);
          for (int e = 0; e < cnt; ++e) {
            const bool fe = flagged(i0 + e);
            if (fe && !fprev) {
              const uint32_t pos = atomicAdd(&s.n_runs, 1u);
              if (pos < static_cast<uint32_t>(kMaxRuns)) s.runs[pos] = i0 + e;
            }
            fprev = fe;
          }
        }
      };
      auto group_loop = [&](auto value_of) {
        uint32_t gmask = 0;
        int step = 0;
        for (int gidx = tid; gidx < ((dbg & 4) ? 0 : n_groups); gidx += kThreads, ++step) {
          const int i0 = 4 * gidx;
          const uint2 lo = *reinterpret_cast<const uint2*>(q8 + i0);
          const uint2 hi = *reinterpret_cast<const uint2*>(q8 + i0 + 4);
          const uint32_t wa = (sh & 2) ? lo.y : lo.x, wb = (sh & 2) ? hi.x : lo.y, wc = (sh & 2) ? hi.y : hi.x;
          const uint32_t p0 = __funnelshift_r(wa, wb, 16 * (sh & 1)), p1 = __funnelshift_r(wb, wc, 16 * (sh & 1));
          const int xs[4] = {static_cast<int16_t>(p0 & 0xffffu), static_cast<int16_t>(p0 >> 16),
                             static_cast<int16_t>(p1 & 0xffffu), static_cast<int16_t>(p1 >> 16)};
          const int cnt = min(4, n - i0);
          if (cnt == 4) {
            float4 v;
            v.x = value_of(xs[0], 0);
            v.y = value_of(xs[1], 1);
            v.z = value_of(xs[2], 2);
            v.w = value_of(xs[3], 3);
            if (dbg & 1) {
              if (v.x == 123.456f) o[i0] = v.y + v.z + v.w;
            } else {
              *reinterpret_cast<float4*>(o + i0) = v;
            }
          } else {       // tail group: what lies past the window in the staging buffer is not a sample
            for (int e = 0; e < cnt; ++e) o[i0 + e] = value_of(xs[e], 0);
          }
          // some sample of this group is an outlier (or, in the tail group, stale data past the window)
          const bool hit = !(dbg & 32) && (__vimax3_s16x2(p0, p1, hi2) != hi2 || __vimin3_s16x2(p0, p1, lo2) != lo2);
          gmask |= (hit ? 1u : 0u) << (step & 31);
          if ((step & 31) == 31) {          // windows beyond 32,768 samples: empty the mask every 32 steps
            note_runs(gmask, step - 31);
            gmask = 0;
          }
        }
        note_runs(gmask, step & ~31);
      };
      if (use_lut) {
        const float* lutb = lut - vmin;              // indexed by the sample value itself
        if (dbg & 16) {
          // experiment: samples 0 and 2 of a group from the table, 1 and 3 from the float64 pipe
          group_loop([&](int xv, int e) { return (e & 1) ? static_cast<float>(quot(2 * xv - med2)) : lutb[xv]; });
        } else {
          group_loop([&](int xv, int) { return lutb[xv]; });
        }
      } else {
        group_loop([&](int xv, int) { return static_cast<float>(quot(2 * xv - med2)); });
      }
      __syncthreads();   // every sample has its plain quotient; the runs overwrite theirs
      const uint32_t n_runs = (dbg & 64) ? 0u : s.n_runs;
      if (n_runs <= static_cast<uint32_t>(kMaxRuns)) {
        for (uint32_t r = tid; r < n_runs; r += kThreads) walk(s.runs[r]);
      } else {             // list overflowed: find the run starts again
        for (int i = tid; i < n; i += kThreads)
          if (flagged(i) && !(i > 0 && flagged(i - 1))) walk(i);
      }
    }
    __syncthreads();   // stage / scratch are reused by the next read
  }
}

// ------------------------------------------------------------------------------------
// poly(A) end detection.  One CTA per read; one warp per 500-sample window.

constexpr int kRes = 500;               // _TRIM_RESOLUTION, riser/preprocess.py:10
constexpr int kPerLane = (kRes + 31) / 32;   // 16
constexpr int kMaxWindows = 512;        // windows per read (256,000 samples; the live path stops near 18,500)
constexpr int kPolyaThreads = 128;      // 4 warps per read: an 18,000-sample prefix has 36 windows, 9 per warp
constexpr int kPolyaWarps = kPolyaThreads / 32;
constexpr int kWinBins = 1024;          // value range a window's shared-memory histogram covers
constexpr int kWinHist = 33 * 32 + 32;  // per-warp histogram storage: 32 lanes x (odd) segment of up to 33 bins

// rank-k (0-based) key among this warp's keys (all < 2^nbits; padding lanes hold 0xffffffff): binary search
// on the value, one warp-wide count #{key < T} per bit (REDUX) -- the "warp-shuffle selection" of the north
// star.  nbits comes from the window's own range, so a 500-sample window costs ~10 steps, not 16.
__device__ __forceinline__ uint32_t warp_select(const uint32_t (&key)[kPerLane], int nbits, uint32_t k) {
  uint32_t prefix = 0;
  for (int bit = nbits - 1; bit >= 0; --bit) {
    const uint32_t t = prefix | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) cnt += (key[j] < t) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (k >= static_cast<uint32_t>(cnt)) prefix = t;      // rank k lies at or above t
  }
  return prefix;
}

// the two middle order statistics (ranks 249 and 250 of 500) -> their sum
__device__ __forceinline__ uint32_t warp_mid_sum(const uint32_t (&key)[kPerLane], int nbits) {
  const uint32_t v1 = warp_select(key, nbits, kRes / 2 - 1);
  int le = 0;
  uint32_t next = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    le += key[j] <= v1 ? 1 : 0;
    if (key[j] > v1) next = min(next, key[j]);      // (padding keys are 0xffffffff: never the minimum of 500 real keys)
  }
  le = __reduce_add_sync(0xffffffffu, le);
  next = __reduce_min_sync(0xffffffffu, next);
  const uint32_t v2 = (le > kRes / 2) ? v1 : next;
  return v1 + v2;
}

__device__ __forceinline__ int bits_of(uint32_t v) { return 32 - __clz(v); }

// Median and MAD of one 500-sample window from a per-warp shared-memory histogram of x - vmin (range < kWinBins,
// which covers every real squiggle window; wider windows take the bit-descent path below): ~3x fewer instructions
// than two 10..11-step bit-descents.  h becomes the inclusive prefix P[u] = #{x - vmin <= u}; the two middle order
// statistics are found by warp-wide searches on P, and those of |2x - 2 median| from the same P (as in the
// normalise kernel): #{|2(x - vmin) - m| <= d} = P[floor((m + d) / 2)] - P[ceil((m - d) / 2) - 1].
__device__ __forceinline__ void warp_hist_stats(const int (&v)[kPerLane], const bool (&ok)[kPerLane], int vmin, int range,
                                                uint32_t* h, int& med2, uint32_t& mad4) {
  const int lane = threadIdx.x & 31;
  const int nb = range + 1;
  int seg = (nb + 31) >> 5;                 // bins per lane; odd, so the lanes' segments start in distinct banks
  seg |= 1;
  const int n_clear = (32 * seg + 3) >> 2;  // 16-byte words to clear
  for (int i = lane; i < n_clear; i += 32) reinterpret_cast<uint4*>(h)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < kPerLane; ++j)
    if (ok[j]) atomicAdd(&h[v[j] - vmin], 1u);
  __syncwarp();
  // in-place inclusive prefix: lane sums its segment, warp scan of the sums, second sweep writes
  uint32_t* mine = h + lane * seg;
  uint32_t tot = 0;
  for (int i = 0; i < seg; ++i) tot += mine[i];
  uint32_t inc = tot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += u;
  }
  uint32_t run = inc - tot;
  for (int i = 0; i < seg; ++i) {
    run += mine[i];
    mine[i] = run;
  }
  __syncwarp();
  // smallest u with P[u] >= want, for want = 250 (lanes 0-15) and 251 (lanes 16-31): 16-ary search
  const int half = lane >> 4, sub = lane & 15;
  const uint32_t want = static_cast<uint32_t>(kRes / 2 + half);
  const uint32_t hmask = half ? 0xffff0000u : 0x0000ffffu;
  int lo = 0, hi = range;                   // invariant: P[hi] >= want (P[range] = 500)
  while (__any_sync(0xffffffffu, lo < hi)) {   // the two half-warp searches run in lock-step; a finished one idles
    const int step = (hi - lo + 16) >> 4;   // ceil(span / 16)
    const int u = min(lo + (sub + 1) * step - 1, hi);
    const uint32_t okm = __ballot_sync(0xffffffffu, h[u] >= want) & hmask;
    const int f = (__ffs(okm) - 1) & 15;    // sub-lane 15 probes hi: never empty
    hi = min(lo + (f + 1) * step - 1, hi);
    lo = lo + f * step;
  }
  const int u1 = __shfl_sync(0xffffffffu, lo, 0), u2 = __shfl_sync(0xffffffffu, lo, 16);
  const int m = u1 + u2;
  med2 = 2 * vmin + m;
  // the same search on the distance d = |2(x - vmin) - m|
  lo = 0;
  hi = max(m, 2 * range - m);               // count(hi) = 500
  while (__any_sync(0xffffffffu, lo < hi)) {   // the two half-warp searches run in lock-step; a finished one idles
    const int step = (hi - lo + 16) >> 4;
    const int d = min(lo + (sub + 1) * step - 1, hi);
    const int hi_u = min((m + d) >> 1, range);
    const int lo_u = (m - d + 1) >> 1;      // ceil((m - d) / 2), may be <= 0
    const uint32_t cnt = h[hi_u] - (lo_u > 0 ? h[lo_u - 1] : 0u);
    const uint32_t okm = __ballot_sync(0xffffffffu, cnt >= want) & hmask;
    const int f = (__ffs(okm) - 1) & 15;
    hi = min(lo + (f + 1) * step - 1, hi);
    lo = lo + f * step;
  }
  mad4 = static_cast<uint32_t>(__shfl_sync(0xffffffffu, lo, 0) + __shfl_sync(0xffffffffu, lo, 16));
  __syncwarp();                             // h is cleared again for the warp's next window
}

__global__ void __launch_bounds__(kPolyaThreads)
polya_kernel(const int16_t* __restrict__ sig, const int64_t* __restrict__ off,
             const int32_t* __restrict__ nsamp, int B, int32_t* __restrict__ polya_end,
             int32_t* __restrict__ polya_start, int32_t* __restrict__ stats, int max_windows) {
  __shared__ int32_t w_sum[kMaxWindows];
  __shared__ int32_t w_mad4[kMaxWindows];
  __shared__ __align__(16) uint32_t w_hist[kPolyaWarps][kWinHist];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int n = nsamp[b];
    const int16_t* g = sig + off[b];
    const int nw = min(n / kRes, kMaxWindows);
    // 4-byte loads when the read is 4-byte aligned (a window is 1,000 bytes): a lane takes sample pairs -- which lane
    // holds which sample is irrelevant to the statistics -- and the NEXT window of this warp is fetched while the
    // current one is worked on
    const bool al4 = (reinterpret_cast<uintptr_t>(g) & 3) == 0;
    uint32_t raw[kPerLane / 2];
    auto fetch = [&](int w) {
      const uint32_t* wp2 = reinterpret_cast<const uint32_t*>(g + w * kRes);
#pragma unroll
      for (int j = 0; j < kPerLane / 2; ++j) {
        const int idx = j * 32 + lane;
        raw[j] = (idx < kRes / 2) ? __ldg(wp2 + idx) : 0u;
      }
    };
    if (al4 && warp < nw) fetch(warp);
    for (int w = warp; w < nw; w += kPolyaWarps) {
      const int16_t* wp = g + w * kRes;
      int v[kPerLane];
      int sum = 0, vmin = 32767, vmax = -32768;
      bool ok[kPerLane];
      if (al4) {
#pragma unroll
        for (int j = 0; j < kPerLane / 2; ++j) {
          ok[2 * j] = ok[2 * j + 1] = (j * 32 + lane) < kRes / 2;
          v[2 * j] = static_cast<int16_t>(raw[j] & 0xffffu);
          v[2 * j + 1] = static_cast<int16_t>(raw[j] >> 16);
        }
        if (w + kPolyaWarps < nw) fetch(w + kPolyaWarps);
      } else {
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
          const int idx = j * 32 + lane;
          ok[j] = idx < kRes;
          v[j] = ok[j] ? static_cast<int>(wp[idx]) : 0;
        }
      }
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) {
        sum += v[j];
        vmin = min(vmin, ok[j] ? v[j] : 32767);
        vmax = max(vmax, ok[j] ? v[j] : -32768);
      }
      sum = __reduce_add_sync(0xffffffffu, sum);
      vmin = __reduce_min_sync(0xffffffffu, vmin);
      vmax = __reduce_max_sync(0xffffffffu, vmax);
      if (vmax - vmin < kWinBins) {
        int med2h;
        uint32_t mad4h;
        warp_hist_stats(v, ok, vmin, vmax - vmin, w_hist[warp], med2h, mad4h);
        if (lane == 0) {
          w_sum[w] = sum;
          w_mad4[w] = static_cast<int32_t>(mad4h);
          if (stats && w < max_windows) {
            int32_t* st = stats + (static_cast<int64_t>(b) * max_windows + w) * 3;
            st[0] = sum;
            st[1] = med2h;
            st[2] = static_cast<int32_t>(mad4h);
          }
        }
        continue;
      }
      // median on keys x - min (as many bits as the window's range needs)
      uint32_t key[kPerLane];
#pragma unroll
      for (int j = 0; j < kPerLane; ++j)
        key[j] = ok[j] ? static_cast<uint32_t>(v[j] - vmin) : 0xffffffffu;
      const int med2 = 2 * vmin + static_cast<int>(warp_mid_sum(key, bits_of(static_cast<uint32_t>(vmax - vmin))));   // 2 * median
      // MAD on keys |2x - 2 median|
      const uint32_t dmax = static_cast<uint32_t>(max(abs(2 * vmin - med2), abs(2 * vmax - med2)));
#pragma unroll
      for (int j = 0; j < kPerLane; ++j)
        key[j] = ok[j] ? static_cast<uint32_t>(abs(2 * v[j] - med2)) : 0xffffffffu;
      const uint32_t mad4 = warp_mid_sum(key, bits_of(dmax));                         // 4 * MAD
      if (lane == 0) {
        w_sum[w] = sum;
        w_mad4[w] = static_cast<int32_t>(mad4);
        if (stats && w < max_windows) {
          int32_t* st = stats + (static_cast<int64_t>(b) * max_windows + w) * 3;
          st[0] = sum;
          st[1] = med2;
          st[2] = static_cast<int32_t>(mad4);
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      // riser/preprocess.py:45-72.  0 doubles as "unset" exactly like Python truthiness.  The per-window tests
      // (three float64 divides each) are independent, so the lanes evaluate 32 windows at a time and the two-state
      // scan runs on the ballots: start = first window > 0 with mean_change > 20 and mad <= 20; end = first window
      // after it with mad > 20.
      int pstart = 0, pend = 0;
      for (int w0 = 0; w0 < nw && pend == 0; w0 += 32) {
        const int w = w0 + lane;
        bool cs = false, ce = false;
        if (w < nw) {
          const double mean = __ddiv_rn(static_cast<double>(w_sum[w]), static_cast<double>(kRes));
          double rolling = mean;
          if (w > 2)
            rolling = __ddiv_rn(static_cast<double>(w_sum[w - 2] + w_sum[w - 1]),
                                static_cast<double>(2 * kRes));
          const double change = __dmul_rn(__ddiv_rn(__dsub_rn(mean, rolling), rolling), 100.0);
          const double mad = static_cast<double>(w_mad4[w]) * 0.25;
          cs = w > 0 && change > 20.0 && mad <= 20.0;
          ce = mad > 20.0;
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, cs);
        uint32_t me = __ballot_sync(0xffffffffu, ce);
        if (pstart == 0 && ms) pstart = (w0 + __ffs(ms) - 1) * kRes;
        if (pstart != 0) {
          const int ws = pstart / kRes - w0;            // start window relative to this chunk (< 0: earlier chunk)
          if (ws > 0) me &= ~((1u << ws) - 1u);
          if (me) pend = (w0 + __ffs(me) - 1) * kRes;
        }
      }
      if (lane == 0) {
        polya_end[b] = pend ? pend : -1;
        if (polya_start) polya_start[b] = pstart ? pstart : -1;
      }
    }
    __syncthreads();
  }
}

// riser/control.py:36-60 length gating
__global__ void select_window_kernel(const int32_t* __restrict__ n, const int32_t* __restrict__ cached,
                                     const int32_t* __restrict__ detected, int B, int min_len, int max_len,
                                     int fixed_trim, int32_t* __restrict__ start, int32_t* __restrict__ len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nb = n[b];
  int end = cached ? cached[b] : -1;
  if (end < 0 && detected) end = detected[b];
  int st = 0, ln = 0;
  if (end > 0) {                               // Python truthiness: 0 / None = not found
    const int avail = nb - (end + 1);
    if (avail >= min_len) {
      st = end + 1;
      ln = min(avail, max_len);
    }
  } else if (nb > fixed_trim + max_len) {      // preprocess.py:84-85 (strict)
    st = fixed_trim;
    ln = max_len;
  }
  start[b] = st;
  len[b] = ln;
}

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_sm_count = 148;
  }
  return g_sm_count;
}

}  // namespace
}  // namespace riser

using namespace riser;

extern "C" int riser_version(void) { return 100; }
extern "C" const char* riser_last_error(void) { return err_buf(); }

extern "C" int riser_device_info(int device, int* sm, int* major, int* minor) {
  RISER_REQUIRE(sm && major && minor, "riser_device_info: null output pointer");
  RISER_CUDA_TRY(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, device));
  RISER_CUDA_TRY(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, device));
  RISER_CUDA_TRY(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, device));
  return RISER_OK;
}

extern "C" int riser_normalise_max_len(void) { return kMaxLen; }

extern "C" int riser_normalise(const int16_t* sig, const int64_t* off, const int32_t* start,
                               const int32_t* len, int B, int max_len, float* out, int64_t ld_out,
                               int32_t* med2_mad4, riser_stream_t stream) {
  RISER_REQUIRE(B >= 0, "riser_normalise: B < 0");
  if (B == 0) return RISER_OK;
  RISER_REQUIRE(sig && off && len && out, "riser_normalise: null pointer");
  RISER_REQUIRE(max_len > 0 && max_len <= kMaxLen, "riser_normalise: max_len %d outside (0, %d]",
                max_len, kMaxLen);
  RISER_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ld_out & 3) == 0 && ld_out >= max_len,
                "riser_normalise: out must be 16-byte aligned, ld_out a multiple of 4 and >= max_len");
  // staging buffers of whole 16-byte blocks: window + up to 7 samples of phase + tail block, 128-byte pitch
  const int buf_samples = (max_len + 16 + 63) & ~63;
  const size_t scratch = (sizeof(SelectScratch) + 127) & ~size_t(127);
  static int nbuf_env = -1;
  if (nbuf_env < 0) {
    const char* e = getenv("RISER_NORM_NBUF");
    nbuf_env = e ? atoi(e) : 0;
  }
  // One staging buffer + L2 prefetch of the next read by default: a second buffer (RISER_NORM_NBUF=2, the next
  // read's copy in flight during this read's work) costs a resident CTA per SM and measured slower.
  int nbuf = (nbuf_env == 2) ? 2 : 1;
  if (scratch + 2 * static_cast<size_t>(nbuf) * buf_samples > 227 * 1024) nbuf = 1;
  const size_t smem = scratch + 2 * static_cast<size_t>(nbuf) * buf_samples;
  static size_t configured[64] = {0};   // per device
  int dev = 0;
  RISER_CUDA_TRY(cudaGetDevice(&dev));
  if (smem > configured[dev & 63]) {
    RISER_CUDA_TRY(cudaFuncSetAttribute(normalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    configured[dev & 63] = smem;
  }
  int per_sm = 0;
  RISER_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, normalise_kernel, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  // timing experiments (profiles/README.md): RISER_NORM_CTAS caps the resident CTAs per SM; RISER_NORM_DBG is a
  // bit mask -- 1 no global stores, 2 no histogram atomics, 4 no normalise loop (all three give WRONG results),
  // 8 per-sample float64 quotients instead of the value look-up table (same results)
  static int ctas_env = -1, dbg = 0;
  if (ctas_env < 0) {
    const char* e = getenv("RISER_NORM_CTAS");
    ctas_env = e ? atoi(e) : 0;
    e = getenv("RISER_NORM_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (ctas_env > 0) per_sm = std::min(per_sm, ctas_env);
  const int grid = std::min(B, sm_count() * per_sm);
  normalise_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(sig, off, start, len, B, out, ld_out,
                                                               med2_mad4, nbuf, buf_samples, max_len, dbg);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_select_window(const int32_t* n, const int32_t* cached_end, const int32_t* detected_end,
                                   int B, int min_len, int max_len, int fixed_trim, int32_t* start,
                                   int32_t* len, riser_stream_t stream) {
  RISER_REQUIRE(B >= 0, "riser_select_window: B < 0");
  if (B == 0) return RISER_OK;
  RISER_REQUIRE(n && start && len, "riser_select_window: null pointer");
  select_window_kernel<<<(B + 255) / 256, 256, 0, as_stream(stream)>>>(n, cached_end, detected_end, B, min_len,
                                                                     max_len, fixed_trim, start, len);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}

extern "C" int riser_polya_end(const int16_t* sig, const int64_t* off, const int32_t* n, int B,
                               int32_t* polya_end, int32_t* polya_start, int32_t* stats, int max_windows,
                               riser_stream_t stream) {
  RISER_REQUIRE(B >= 0, "riser_polya_end: B < 0");
  if (B == 0) return RISER_OK;
  RISER_REQUIRE(sig && off && n && polya_end, "riser_polya_end: null pointer");
  RISER_REQUIRE(!stats || max_windows > 0, "riser_polya_end: stats given but max_windows <= 0");
  int per_sm = 0;
  RISER_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, polya_kernel, kPolyaThreads, 0));
  const int grid = std::min(B, sm_count() * std::max(per_sm, 1));
  polya_kernel<<<grid, kPolyaThreads, 0, as_stream(stream)>>>(sig, off, n, B, polya_end, polya_start, stats, max_windows);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}
