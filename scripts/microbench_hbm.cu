// Stand-alone ceiling probe for the basis-streaming access pattern (sm_100a):
//   mode 0  flat read, LDG.128 x 8 per thread and iteration, grid-stride
//   mode 1  the kernels' pattern: persistent blocks, each owns a column range; for every tile and every group of R
//           rows one producer thread issues R bulk copies of SEG bytes (rows LD bytes apart) into a ring of NS
//           stages; 8 consumer warps read the stage with LDS.128 and add it up
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench_hbm scripts/microbench_hbm.cu
// Run:    scripts/microbench_hbm            (prints one line per configuration; every buffer > L2)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../experiments_lanczos_adjoints_b200/csrc/tma_pipeline.cuh"

using namespace bl;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

__global__ void __launch_bounds__(256) k_flat(const float4* __restrict__ p, size_t nvec, float* out) {
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < nvec; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcs(p + i + k * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  for (; i < nvec; i += stride) {
    float4 v = __ldcs(p + i);
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 123.456f) out[0] = s;
}

// rows: m rows, ld floats apart; n columns.  Block b owns columns [b*per, (b+1)*per); tiles of SEG/4 floats.
template <int R>
__global__ void __launch_bounds__(288) k_ring(const float* __restrict__ base, long long ld, int m, long long n,
                                              int seg_floats, int stages, int pf, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* st = reinterpret_cast<float*>(smem);  // [stages][R][seg_floats]
  uint64_t* full = reinterpret_cast<uint64_t*>(st + (size_t)stages * R * seg_floats);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, 8);
    }
    tma::fence_barrier_init();
  }
  __syncthreads();
  long long per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + 31) / 32 * 32;
  const long long c0 = per * blockIdx.x, c1 = c0 + per < n ? c0 + per : n;
  const int ntiles = c0 < c1 ? (int)((c1 - c0 + seg_floats - 1) / seg_floats) : 0;
  const int ngroups = (m + R - 1) / R;
  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      const int total = ntiles * ngroups;
      auto prefetch = [&](int k) {  // L2 prefetch of the k-th (tile, group) of this block
        if (k >= total) return;
        const int t = k / ngroups, g = k % ngroups;
        const long long tc0 = c0 + (long long)t * seg_floats;
        const int len = (int)((c1 - tc0) < seg_floats ? (c1 - tc0) : seg_floats);
        const int rows_here = m - g * R < R ? m - g * R : R;
        for (int r = 0; r < rows_here; ++r)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (long long)(g * R + r) * ld + tc0),
                       "r"((uint32_t)len * 4u)
                       : "memory");
      };
      if (pf > 0)
        for (int k = stages; k < stages + pf; ++k) prefetch(k);
      for (int t = 0; t < ntiles; ++t) {
        const long long tc0 = c0 + (long long)t * seg_floats;
        const int len = (int)((c1 - tc0) < seg_floats ? (c1 - tc0) : seg_floats);
        const uint32_t bytes = (uint32_t)len * 4u;
        for (int g = 0; g < ngroups; ++g, ++it) {
          const int rows_here = m - g * R < R ? m - g * R : R;
          const int s = it % stages;
          if (pf > 0) prefetch(it + stages + pf);
          tma::mbar_wait(empty + s, ((it / stages) & 1) ^ 1);
          tma::mbar_arrive_expect_tx(full + s, bytes * rows_here);
          float* dst = st + (size_t)s * R * seg_floats;
          for (int r = 0; r < rows_here; ++r)
            tma::bulk_g2s(dst + (size_t)r * seg_floats, base + (long long)(g * R + r) * ld + tc0, bytes, full + s);
        }
      }
    }
  } else {
    float acc = 0.f;
    int it = 0;
    for (int t = 0; t < ntiles; ++t) {
      const long long tc0 = c0 + (long long)t * seg_floats;
      const int len = (int)((c1 - tc0) < seg_floats ? (c1 - tc0) : seg_floats);
      for (int g = 0; g < ngroups; ++g, ++it) {
        const int s = it % stages;
        tma::mbar_wait(full + s, (it / stages) & 1);
        const int rows_here = m - g * R < R ? m - g * R : R;
        // 8 warps share the R rows of the stage: warp w reads rows w, w+8, ... (R >= 8) or a slice of a row (R < 8)
        if (R >= 8) {
          for (int r = warp; r < rows_here; r += 8) {
            const float4* row = reinterpret_cast<const float4*>(st + ((size_t)s * R + r) * seg_floats);
            for (int u = lane; u * 4 < len; u += 32) {
              const float4 v = row[u];
              acc += v.x + v.y + v.z + v.w;
            }
          }
        } else {
          const int per_row = 8 / R;  // warps per row
          const int r = warp / per_row, part = warp % per_row;
          if (r < rows_here) {
            const float4* row = reinterpret_cast<const float4*>(st + ((size_t)s * R + r) * seg_floats);
            for (int u = lane + 32 * part; u * 4 < len; u += 32 * per_row) {
              const float4 v = row[u];
              acc += v.x + v.y + v.z + v.w;
            }
          }
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(empty + s);
      }
    }
    if (acc == 123.456f) out[0] = acc;
  }
}

template <int R>
float run_ring(const float* base, long long ld, int m, long long n, int seg_floats, int stages, int pf, int grid,
               float* out, int reps) {
  const size_t smem = (size_t)stages * R * seg_floats * 4 + 2 * stages * 8 + 64;
  CK(cudaFuncSetAttribute(k_ring<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_ring<R><<<grid, 288, smem>>>(base, ld, m, n, seg_floats, stages, pf, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  const long long n = 1000000, ld = 1000064;  // the headline's row length and padded stride
  const int M = 400;                          // 1.6 GB: every pass misses L2
  float *buf, *out;
  CK(cudaMalloc(&buf, (size_t)M * ld * 4));
  CK(cudaMemset(buf, 0, (size_t)M * ld * 4));
  CK(cudaMalloc(&out, 16));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  {  // flat read
    const size_t nvec = (size_t)M * ld / 4;
    for (int bps : {4, 8, 16}) {
      for (int i = 0; i < 2; ++i) k_flat<<<sms * bps, 256>>>(reinterpret_cast<float4*>(buf), nvec, out);
      CK(cudaEventRecord(e0));
      const int reps = 5;
      for (int i = 0; i < reps; ++i) k_flat<<<sms * bps, 256>>>(reinterpret_cast<float4*>(buf), nvec, out);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::printf("flat LDG.128x8  blocks/SM %2d : %7.1f GB/s\n", bps, (double)nvec * 16 / (ms / reps) / 1e6);
    }
  }
  {  // cudaMemcpy D2D for reference (read + write)
    const size_t bytes = (size_t)(M / 2) * ld * 4;
    CK(cudaMemcpy(buf + (size_t)(M / 2) * ld, buf, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(buf + (size_t)(M / 2) * ld, buf, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::printf("cudaMemcpy D2D (read+write bytes)  : %7.1f GB/s\n", 2.0 * bytes / (ms / 5) / 1e6);
  }
  struct Cfg {
    int R, seg_floats, stages, bps, pf;
  };
  const std::vector<Cfg> cfgs = {
      {8, 1024, 3, 2, 0},   // the kernels today: 96 KB ring, 2 blocks/SM
      {8, 848, 3, 2, 0},    // balanced tiles for per = 3392
      {4, 1696, 3, 2, 0},   // two tiles per block
      {4, 1024, 6, 2, 0},   // finer stages, same ring
      {2, 1024, 12, 2, 0}, {4, 848, 6, 2, 0}, {4, 848, 7, 2, 0}, {2, 1696, 7, 2, 0}, {1, 1696, 14, 2, 0}, {1, 3392, 7, 2, 0},
      {8, 1024, 3, 2, 2},   // + L2 prefetch this many stages beyond the ring
      {8, 1024, 3, 2, 4}, {8, 1024, 3, 2, 8}, {8, 848, 3, 2, 4}, {8, 848, 3, 2, 8}, {4, 1696, 3, 2, 4}, {4, 1696, 3, 2, 8},
      {4, 848, 6, 2, 8}, {4, 848, 6, 2, 16}, {8, 848, 2, 2, 6}, {8, 848, 3, 2, 13}, {8, 848, 3, 2, 26},
      {8, 848, 3, 1, 8}, {8, 1024, 6, 1, 8},
  };
  for (int m : {96, 48, 16}) {
    for (const Cfg& c : cfgs) {
      const int grid = sms * c.bps;
      const int passes = M / m;  // distinct row sets so that nothing is served from L2
      float ms = 0.f;
      int reps = 0;
      // walk the buffer: each launch reads rows [p*m, (p+1)*m)
      auto go = [&](int p) {
        const float* base = buf + (size_t)p * m * ld;
        switch (c.R) {
          case 1: return run_ring<1>(base, ld, m, n, c.seg_floats, c.stages, c.pf, grid, out, 1);
          case 2: return run_ring<2>(base, ld, m, n, c.seg_floats, c.stages, c.pf, grid, out, 1);
          case 4: return run_ring<4>(base, ld, m, n, c.seg_floats, c.stages, c.pf, grid, out, 1);
          case 8: return run_ring<8>(base, ld, m, n, c.seg_floats, c.stages, c.pf, grid, out, 1);
          default: return run_ring<16>(base, ld, m, n, c.seg_floats, c.stages, c.pf, grid, out, 1);
        }
      };
      go(0);  // warm-up (module load, attribute); its rows are evicted again by the passes that follow
      for (int p = 1; p <= passes; ++p) {
        ms += go(p % passes);
        ++reps;
      }
      ms /= reps;
      std::printf("ring m=%3d R=%2d seg=%5d B stages=%2d blocks/SM=%d ring=%3d KB L2-prefetch=%2d : %7.1f us  %7.1f GB/s\n", m,
                  c.R, c.seg_floats * 4, c.stages, c.bps, c.R * c.seg_floats * 4 * c.stages / 1024, c.pf, ms * 1e3,
                  (double)m * n * 4 / ms / 1e6);
    }
  }
  return 0;
}
