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

namespace riser {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kBins = 2048;             // first-pass histogram bins
constexpr int kBinsPerThread = kBins / kThreads;
constexpr int kSubBins = 128;           // refinement pass (shift <= 7)
constexpr int kMaxLen = 98304;          // samples staged in smem (192 KB)

constexpr double kOutlierLimit = 3.5;   // riser/preprocess.py:6
constexpr double kScalingFactor = 1.4826;  // riser/preprocess.py:7

struct SelectScratch {
  uint32_t hist[kBins];
  uint32_t sub[kSubBins];
  uint32_t warp_sums[kWarps];
  uint32_t res[4];      // bin1, before1, bin2, before2
  uint32_t refined;
  int32_t red[2 * kWarps];
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

// Exact order statistics k1 <= k2 (0-based) of the n keys key(0..n-1), all <= maxkey.
// Pass A: histogram of key >> shift with shift chosen so that it fits kBins; block scan
// locates the bins holding the two ranks.  Pass B (only when shift > 0): histogram of the
// low bits inside the located bin.  Block-uniform control flow; all threads get r1, r2.
template <class KeyFn>
__device__ int select_two(KeyFn key, int n, uint32_t maxkey, uint32_t k1, uint32_t k2,
                          SelectScratch& s, uint32_t& r1, uint32_t& r2, bool keep_prefix = false) {
  const int tid = threadIdx.x;
  int shift = 0;
  while ((maxkey >> shift) >= static_cast<uint32_t>(kBins)) ++shift;
  for (int i = tid; i < kBins; i += kThreads) s.hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) atomicAdd(&s.hist[key(i) >> shift], 1u);
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
    for (int i = tid; i < n; i += kThreads) {
      const uint32_t kx = key(i);
      if ((kx >> shift) == bin) atomicAdd(&s.sub[kx & mask], 1u);
    }
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

__device__ __forceinline__ double norm_value(int x, double median, double denom) {
  // (x - median) / (1.4826 * mad), riser/preprocess.py:122-125, IEEE round-to-nearest
  return __ddiv_rn(__dsub_rn(static_cast<double>(x), median), denom);
}
__device__ __forceinline__ double clip_outlier(double v) {
  // riser/preprocess.py:141-147
  if (v > kOutlierLimit) return kOutlierLimit;
  if (v < -kOutlierLimit) return -kOutlierLimit;
  return v;
}

__global__ void __launch_bounds__(kThreads)
normalise_kernel(const int16_t* __restrict__ sig, const int64_t* __restrict__ off,
                 const int32_t* __restrict__ start, const int32_t* __restrict__ len, int B,
                 float* __restrict__ out, int64_t ld_out, int32_t* __restrict__ med2_mad4) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelectScratch& s = *reinterpret_cast<SelectScratch*>(smem_raw);
  int16_t* stage = reinterpret_cast<int16_t*>(smem_raw + ((sizeof(SelectScratch) + 15) & ~size_t(15)));
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int n = len[b];
    if (n <= 0) continue;
    const int16_t* g = sig + off[b] + (start ? start[b] : 0);
    float* o = out + static_cast<int64_t>(b) * ld_out;

    // ---- stage the window in smem at the same 16-byte phase as in global memory
    const int a = static_cast<int>((reinterpret_cast<uintptr_t>(g) >> 1) & 7);
    const int16_t* x = stage + a;                    // x[i] == g[i]
    const int n_chunks = (a + n + 7) >> 3;
    const uint4* g4 = reinterpret_cast<const uint4*>(g - a);
    uint4* s4 = reinterpret_cast<uint4*>(stage);
    int vmin = 32767, vmax = -32768;
    auto take16 = [&](const uint4& v) {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e0 = static_cast<int16_t>(w[j] & 0xffff);
        const int e1 = static_cast<int16_t>(w[j] >> 16);
        vmin = min(vmin, min(e0, e1));
        vmax = max(vmax, max(e0, e1));
      }
    };
    // interior chunks (fully inside the window) are 16-byte loads, four in flight per thread;
    // the (at most two) edge chunks are peeled into scalar loads
    const int c_lo = (a > 0) ? 1 : 0;
    const int c_hi = (a + n) >> 3;          // chunks [c_lo, c_hi) are interior
    for (int c = c_lo + tid; c < c_hi; c += 4 * kThreads) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u * kThreads < c_hi) v[u] = __ldg(g4 + c + u * kThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u * kThreads < c_hi) {
          s4[c + u * kThreads] = v[u];
          take16(v[u]);
        }
    }
    if (tid < 2) {
      const int c = tid ? c_hi : 0;
      if ((tid == 0 && c_lo == 1) || (tid == 1 && c_hi < n_chunks && !(c_hi == 0 && c_lo == 1))) {
        const int lo = 8 * c - a;
        for (int i = max(lo, 0); i < min(lo + 8, n); ++i) {
          const int e = g[i];
          stage[a + i] = static_cast<int16_t>(e);
          vmin = min(vmin, e);
          vmax = max(vmax, e);
        }
      }
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
    const int shift_used = select_two([&](int i) { return static_cast<uint32_t>(x[i] - vmin); }, n,
                                      static_cast<uint32_t>(range), k1, k2, s, r1, r2, true);
    const int med2 = 2 * vmin + static_cast<int>(r1 + r2);

    // ---- MAD on keys d = |2x - med2| = 2|x - median| -> mad4 = 4 * MAD (exact)
    const uint32_t dmax = static_cast<uint32_t>(max(abs(2 * vmin - med2), abs(2 * vmax - med2)));
    uint32_t mad4;
    if (shift_used == 0) {
      // The value histogram already holds the whole distribution: with P[u] = #{x - vmin <= u}
      // and m = med2 - 2*vmin, #{|2x - med2| <= d} = P[floor((m+d)/2)] - P[ceil((m-d)/2) - 1];
      // the two middle order statistics of d are found by bisection on d (threads 0 and 1).
      if (tid < 2) {
        const int m = static_cast<int>(r1 + r2);
        const uint32_t want = (tid ? k2 : k1) + 1;
        int lo_d = 0, hi_d = static_cast<int>(dmax);
        while (lo_d < hi_d) {
          const int d = (lo_d + hi_d) >> 1;
          int hi_u = (m + d) >> 1;
          hi_u = min(hi_u, range);
          const int lo_u = (m - d + 1) >> 1;                 // ceil((m - d) / 2), may be <= 0
          const uint32_t cnt = s.hist[hi_u] - (lo_u > 0 ? s.hist[lo_u - 1] : 0u);
          if (cnt >= want) hi_d = d; else lo_d = d + 1;
        }
        s.res[tid] = static_cast<uint32_t>(lo_d);
      }
      __syncthreads();
      mad4 = s.res[0] + s.res[1];
      __syncthreads();
    } else {
      select_two([&](int i) { return static_cast<uint32_t>(abs(2 * x[i] - med2)); }, n, dmax, k1, k2,
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
      __syncthreads();
      continue;
    }

    const double median = static_cast<double>(med2) * 0.5;
    const double mad = static_cast<double>(mad4) * 0.25;
    const double denom = __dmul_rn(kScalingFactor, mad);

    // ---- outlier threshold in the integer key domain: |z| > 3.5  <=>  d > dthr, where
    //      z = fl((d/2) / denom) is monotone in d.  Found once per read with exact divides.
    if (tid == 0) {
      auto zval = [&](uint32_t d) { return __ddiv_rn(static_cast<double>(d) * 0.5, denom); };
      uint32_t d0 = static_cast<uint32_t>(__dmul_rn(7.0, denom));
      while (zval(d0 + 1) <= kOutlierLimit) ++d0;
      while (d0 > 0 && zval(d0) > kOutlierLimit) --d0;
      s.refined = d0;
    }
    __syncthreads();
    const uint32_t dthr = s.refined;
    auto flagged = [&](int i) { return static_cast<uint32_t>(abs(2 * x[i] - med2)) > dthr; };
    // Bulk path: (x - median) / denom == k / D with k = 2x - 2*median (an exact small integer)
    // and D = 2 * denom (exact).  With y = RN(1/D), q0 = RN(k*y), e = k - q0*D (exact in one
    // FMA), RN(q0 + e*y) is the correctly rounded quotient (Markstein's final division step),
    // i.e. bit-identical to the reference's float64 divide at 3 FP64 ops instead of a full DDIV.
    const double Dd = __dmul_rn(2.0, denom);
    const double yrcp = __ddiv_rn(1.0, Dd);
    auto fast_k = [&](int ki) {
      const double k = static_cast<double>(ki);
      const double q0 = __dmul_rn(k, yrcp);
      const double e = __fma_rn(-q0, Dd, k);
      return static_cast<float>(__fma_rn(e, yrcp, q0));
    };
    auto fast_norm = [&](int xv) { return fast_k(2 * xv - med2); };

    // ---- normalise + smooth, 4 samples per thread-iteration, 16-byte stores
    for (int gidx = tid; gidx < n_groups; gidx += kThreads) {
      const int i0 = 4 * gidx;
      if (i0 + 4 <= n) {
        // common case: four in-range samples, none an outlier -> straight-line code
        const int k0 = 2 * x[i0] - med2, k1 = 2 * x[i0 + 1] - med2, k2 = 2 * x[i0 + 2] - med2,
                  k3 = 2 * x[i0 + 3] - med2;
        const uint32_t kmax = static_cast<uint32_t>(max(max(abs(k0), abs(k1)), max(abs(k2), abs(k3))));
        if (kmax <= dthr) {
          float4 v;
          v.x = fast_k(k0);
          v.y = fast_k(k1);
          v.z = fast_k(k2);
          v.w = fast_k(k3);
          *reinterpret_cast<float4*>(o + i0) = v;
          continue;
        }
      }
      const int cnt = min(4, n - i0);
      bool f[4];
      bool any = false;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f[e] = (e < cnt) && flagged(i0 + e);
        any |= f[e];
      }
      if (!any && cnt == 4) {   // (kept for the slow path's bookkeeping; the common case returned above)
        float4 v;
        v.x = fast_norm(x[i0]);
        v.y = fast_norm(x[i0 + 1]);
        v.z = fast_norm(x[i0 + 2]);
        v.w = fast_norm(x[i0 + 3]);
        *reinterpret_cast<float4*>(o + i0) = v;
        continue;
      }
      for (int e = 0; e < cnt; ++e) {
        const int i = i0 + e;
        if (!f[e]) {
          o[i] = static_cast<float>(norm_value(x[i], median, denom));
          continue;
        }
        const bool prev_flag = (i > 0) && (e > 0 ? f[e - 1] : flagged(i - 1));
        if (prev_flag) continue;   // inside a run: the thread at the run start walks it
        // Walk the run of consecutive outliers starting at i, sequentially, exactly as
        // riser/preprocess.py:130-138: arr[j-1] is already updated, arr[j+1] is still raw.
        double prev = (i > 0) ? norm_value(x[i - 1], median, denom) : 0.0;
        for (int j = i; j < n && flagged(j); ++j) {
          double nv;
          if (j == 0) {
            nv = (n > 1) ? norm_value(x[1], median, denom) : norm_value(x[0], median, denom);
          } else if (j == n - 1) {
            nv = prev;
          } else {
            nv = clip_outlier(__dmul_rn(__dadd_rn(prev, norm_value(x[j + 1], median, denom)), 0.5));
          }
          o[j] = static_cast<float>(nv);
          prev = nv;
        }
      }
    }
    __syncthreads();   // stage / scratch are reused by the next read
  }
}

// ------------------------------------------------------------------------------------
// poly(A) end detection.  One CTA per read; one warp per 500-sample window.

constexpr int kRes = 500;               // _TRIM_RESOLUTION, riser/preprocess.py:10
constexpr int kPerLane = (kRes + 31) / 32;   // 16
constexpr int kMaxWindows = 2048;

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

__global__ void __launch_bounds__(kThreads)
polya_kernel(const int16_t* __restrict__ sig, const int64_t* __restrict__ off,
             const int32_t* __restrict__ nsamp, int B, int32_t* __restrict__ polya_end,
             int32_t* __restrict__ polya_start, int32_t* __restrict__ stats, int max_windows) {
  __shared__ int32_t w_sum[kMaxWindows];
  __shared__ int32_t w_mad4[kMaxWindows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int n = nsamp[b];
    const int16_t* g = sig + off[b];
    const int nw = min(n / kRes, kMaxWindows);
    for (int w = warp; w < nw; w += kWarps) {
      const int16_t* wp = g + w * kRes;
      int v[kPerLane];
      int sum = 0, vmin = 32767, vmax = -32768;
      bool ok[kPerLane];
      if ((reinterpret_cast<uintptr_t>(wp) & 3) == 0) {      // 4-byte loads: lane takes sample pairs (which lane holds
        const uint32_t* wp2 = reinterpret_cast<const uint32_t*>(wp);   // which sample is irrelevant to the statistics)
#pragma unroll
        for (int j = 0; j < kPerLane / 2; ++j) {
          const int idx = j * 32 + lane;
          ok[2 * j] = ok[2 * j + 1] = idx < kRes / 2;
          const uint32_t pr = ok[2 * j] ? __ldg(wp2 + idx) : 0u;
          v[2 * j] = static_cast<int16_t>(pr & 0xffffu);
          v[2 * j + 1] = static_cast<int16_t>(pr >> 16);
        }
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
    if (threadIdx.x == 0) {
      // riser/preprocess.py:45-72.  0 doubles as "unset" exactly like Python truthiness.
      int pstart = 0, pend = 0;
      for (int w = 0; w < nw; ++w) {
        const int i = w * kRes;
        const double mean = __ddiv_rn(static_cast<double>(w_sum[w]), static_cast<double>(kRes));
        double rolling = mean;
        if (i > 2 * kRes)
          rolling = __ddiv_rn(static_cast<double>(w_sum[w - 2] + w_sum[w - 1]),
                              static_cast<double>(2 * kRes));
        const double change = __dmul_rn(__ddiv_rn(__dsub_rn(mean, rolling), rolling), 100.0);
        const double mad = static_cast<double>(w_mad4[w]) * 0.25;
        if (pstart == 0 && change > 20.0 && mad <= 20.0) pstart = i;
        if (pstart != 0 && pend == 0 && mad > 20.0) pend = i;
      }
      polya_end[b] = pend ? pend : -1;
      if (polya_start) polya_start[b] = pstart ? pstart : -1;
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
  const size_t smem = ((sizeof(SelectScratch) + 15) & ~size_t(15)) + 2 * (static_cast<size_t>(max_len) + 16);
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
  const int grid = std::min(B, sm_count() * per_sm);
  normalise_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(sig, off, start, len, B, out, ld_out,
                                                               med2_mad4);
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
  const int grid = std::min(B, sm_count() * 4);
  polya_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(sig, off, n, B, polya_end, polya_start, stats, max_windows);
  RISER_CUDA_TRY(cudaGetLastError());
  return RISER_OK;
}
