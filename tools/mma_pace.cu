// Microbenchmark: what paces a tcgen05.mma on B200 (sm_100a)?  One CTA per SM issues a long train of MMAs from one
// thread -- operands in shared memory (SS) or A in tensor memory (TS), K-major SWIZZLE_64B (32-channel K blocks)
// or SWIZZLE_128B (64-channel), kind::f16 or kind::f8f6f4, N from 32 to 256, into one or several accumulators --
// and reports cycles per MMA; then the drain rate of tcgen05.ld with 4 / 8 / 16 warps.  Timing only (operand data is
// whatever shared memory holds).  Design input for csrc/convnet.cu and csrc/resnet_tc.cu (profiles/README.md).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build_ab/mma_pace tools/mma_pace.cu && build_ab/mma_pace
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../riser_b200/csrc/common.cuh"

using namespace riser;

struct Cfg {
  int n, sw128, f8, accs, ts, reps, grid;
};

__device__ __forceinline__ uint64_t desc(uint32_t addr, bool sw128) {
  const uint64_t hi = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>((sw128 ? 1024 : 512) >> 4) << 32) |
                      (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(sw128 ? 2 : 4) << 61);
  return hi | static_cast<uint64_t>(addr >> 4);
}

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_kernel(Cfg c, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < (136 * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tbase, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base + 136 * 128);
    const uint32_t idesc = umma_idesc_f16(128, c.n);
    const uint64_t da = desc(a_addr, c.sw128), db = desc(b_addr, c.sw128);
    const uint32_t acc_stride = (c.n + 31) & ~31;
    const uint32_t a_tmem = tbase + 448;                   // TS: A operand columns (8 per K step of fp16)
    // (no division or table look-up in the issue loop: the issuing thread's own instruction stream must not be what
    //  is measured -- a first version with two `%` per iteration read 257 cycles per MMA whatever N was)
    const uint32_t d0 = tbase, d1 = tbase + (c.accs > 1 ? acc_stride : 0);
    const uint64_t da1 = da + 2, db1 = db + 2;
    long long t0 = clock64();
    if (c.ts) {
      for (int r = 0; r < c.reps; r += 4) {
        umma_f16_ts(d0, a_tmem, db, idesc, 1);
        umma_f16_ts(d1, a_tmem + 8, db1, idesc, 1);
        umma_f16_ts(d0, a_tmem, db, idesc, 1);
        umma_f16_ts(d1, a_tmem + 8, db1, idesc, 1);
      }
    } else if (c.f8) {
      for (int r = 0; r < c.reps; r += 4) {
        umma_f8(d0, da, db, idesc, 1);
        umma_f8(d1, da1, db1, idesc, 1);
        umma_f8(d0, da, db, idesc, 1);
        umma_f8(d1, da1, db1, idesc, 1);
      }
    } else {
      for (int r = 0; r < c.reps; r += 4) {
        umma_f16(d0, da, db, idesc, 1);
        umma_f16(d1, da1, db1, idesc, 1);
        umma_f16(d0, da, db, idesc, 1);
        umma_f16(d1, da1, db1, idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

// drain: `warps` warps (4 per lane quadrant set) read `cols` columns x reps with tcgen05.ld 32x32b.x16
__global__ void __launch_bounds__(512, 1) drain_kernel(int cols, int reps, long long* out) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) {
    tmem_alloc(&tbase, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t t_lane = tbase + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
  const int set = warp >> 2, sets = blockDim.x >> 7;
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (int c16 = 16 * set; c16 < cols; c16 += 16 * sets) {
      uint32_t v[16];
      tmem_ld_32x16(t_lane + c16, v);
      tmem_ld_wait();
      sink += v[0] ^ v[15];
    }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (sink == 0x12345678u) out[1] = sink;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  const size_t smem = 1024 + 136 * 128 + 256 * 128;
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  printf("# cycles per tcgen05.mma (M = 128, one K step), one issuing thread, grid = 148 CTAs; issue / complete\n");
  printf("%-5s %-5s %-4s %-4s %-3s %10s %10s\n", "kind", "sw", "N", "accs", "ts", "issue", "complete");
  const int ns[] = {32, 48, 64, 96, 128, 160, 192, 240, 256};
  for (int f8 = 0; f8 < 2; ++f8)
    for (int sw128 = 0; sw128 < 2; ++sw128)
      for (int n : ns)
        for (int accs : {1, 2}) {
          if (accs * ((n + 31) & ~31) > 448) continue;
          Cfg c{n, sw128, f8, accs, 0, 4000, 148};
          mma_kernel<<<c.grid, 128, smem>>>(c, out);
          if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 1;
          }
          printf("%-5s %-5s %-4d %-4d %-3d %10.1f %10.1f\n", f8 ? "f8" : "f16", sw128 ? "128B" : "64B", n, accs, 0,
                 out[0] / 4000.0, out[1] / 4000.0);
        }
  for (int n : ns) {   // A operand from tensor memory
    Cfg c{n, 0, 0, 1, 1, 4000, 148};
    mma_kernel<<<c.grid, 128, smem>>>(c, out);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("TS launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    printf("%-5s %-5s %-4d %-4d %-3d %10.1f %10.1f\n", "f16", "64B", n, 1, 1, out[0] / 4000.0, out[1] / 4000.0);
  }
  printf("# tcgen05.ld 32x32b.x16 drain: bytes per cycle per SM (128 lanes x cols x 4 B x reps / cycles)\n");
  for (int warps : {4, 8, 16})
    for (int cols : {64, 256}) {
      drain_kernel<<<148, warps * 32>>>(cols, 200, out);
      if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("drain launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
      }
      printf("warps %2d cols %3d: %.1f B/clk\n", warps, cols, 128.0 * cols * 4 * 200 / out[0]);
    }
  return 0;
}
