// TMA-staged versions of the basis-streaming kernels (sm_100a).
//
//   k_dots_tma    : red[j] = <row_j, x>                      pass A of CGS2, adjoint re-projection
//   k_combine_tma : out = s * (sum_k a_k vec_k + sum_j c_j row_j) (+ ||out||^2)   pass C, adjoint back-substitution
//   k_project_tma : x' = x - sum_j c_j row_j  AND  red[j] = <row_j, x'> from ONE read of the rows   (pass B)
//
// Persistent blocks (2 per SM): block b owns a contiguous, balanced range of columns and walks it
// in tiles.  Warp 8 lane 0 is the producer: for every tile and every group of 8 basis rows it
// issues one bulk copy per row segment into a shared-memory stage; warps 0-7 consume.  In
// k_project_tma the whole [m x TILE] tile stays resident so the second Gram-Schmidt projection
// reads the rows from shared memory instead of HBM.
//
// Contract: row buffers are zero-padded up to `ld` (ld*sizeof(T) % 16 == 0), so the 16-byte
// vector that straddles `n` can be copied and multiplied without masking.
#pragma once

#include "krylov_kernels.cuh"
#include "tma_pipeline.cuh"

namespace bl {

constexpr int kGroup = 8;            // basis rows per pipeline stage (= consumer warps)
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = 32 * kConsumerWarps;
constexpr int kStreamThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kStages = 3;

struct RowSource {  // rows [0, n0) from block 0, [n0, n0 + n1) from block 1
  const char* base0 = nullptr;
  long long ldb0 = 0;  // bytes
  int n0 = 0;
  const char* base1 = nullptr;
  long long ldb1 = 0;
  int n1 = 0;
  __device__ __forceinline__ const char* row(int j) const {
    return j < n0 ? base0 + (long long)j * ldb0 : base1 + (long long)(j - n0) * ldb1;
  }
};

struct ColumnRange {
  long long c0, c1;  // [c0, c1) columns of this block, in elements; c1 - c0 multiple of the vector width
  int ntiles;
};

template <typename T>
__device__ __forceinline__ ColumnRange block_columns(long long n, int tile) {
  constexpr int VN = Vec<T>::N;
  const long long n_v = (n + VN - 1) / VN * VN;
  long long per = (n_v + gridDim.x - 1) / gridDim.x;
  per = (per + 31) / 32 * 32;
  ColumnRange r;
  r.c0 = per * blockIdx.x;
  r.c1 = r.c0 + per < n_v ? r.c0 + per : n_v;
  r.ntiles = r.c0 < r.c1 ? (int)((r.c1 - r.c0 + tile - 1) / tile) : 0;
  return r;
}

// Fixed-order cross-block reduction + epilogue, run by the last block (all threads call it).
// Latency-bound (one block reads gridDim.x * nrows doubles from L2), so every lane issues all of
// its loads for two rows before the first add: <= 10 independent 16-byte loads in flight.
template <typename T>
__device__ __forceinline__ void reduce_partials_and_epilogue(int nrows, const double* __restrict__ partials,
                                                             const Epi& epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int G = gridDim.x;
  if ((G & 1) == 0 && G <= 320) {
    const int nv = G >> 1;  // 16-byte vectors per row
    for (int j = warp; j < nrows; j += 2 * nwarps) {
      const int j2 = j + nwarps < nrows ? j + nwarps : j;  // second row (or the same one again)
      const double2* p0 = reinterpret_cast<const double2*>(partials + (size_t)j * G);
      const double2* p1 = reinterpret_cast<const double2*>(partials + (size_t)j2 * G);
      double2 v0[5], v1[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int i = lane + 32 * k;
        const int ic = i < nv ? i : 0;
        v0[k] = __ldcg(p0 + ic);
        v1[k] = __ldcg(p1 + ic);
        if (i >= nv) v0[k] = v1[k] = make_double2(0.0, 0.0);
      }
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        s0 += v0[k].x + v0[k].y;
        s1 += v1[k].x + v1[k].y;
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        epi.red[j] = s0;
        if (j2 != j) epi.red[j2] = s1;
      }
    }
  } else {
    for (int j = warp; j < nrows; j += nwarps) {
      const double* p = partials + (size_t)j * G;
      double s = 0.0;
      for (int b = lane; b < G; b += 32) s += __ldcg(p + b);
      s = warp_sum(s);
      if (lane == 0) epi.red[j] = s;
    }
  }
  __syncthreads();
  run_epilogue<T>(epi);
}

// ---------------------------------------------------------------------------------------------
template <typename T, int TILE>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_dots_tma(RowSource src, int nrows, const T* __restrict__ x, long long n, double* __restrict__ partials,
           unsigned int* counter, Epi epi, int reverse) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int XV = TILE / (32 * VN);  // x vectors per lane
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stages = reinterpret_cast<T*>(smem_raw);  // [kStages][kGroup][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kGroup * TILE * sizeof(T));
  uint64_t* empty = full + kStages;
  double* acc_s = reinterpret_cast<double*>(empty + kStages);  // [nrows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, kConsumerWarps);
    }
    tma::fence_barrier_init();
  }
  for (int j = threadIdx.x; j < nrows; j += blockDim.x) acc_s[j] = 0.0;
  __syncthreads();
  tma::griddep_launch_dependents();

  const ColumnRange cr = block_columns<T>(n, TILE);
  const int ngroups = (nrows + kGroup - 1) / kGroup;

  if (warp == kConsumerWarps) {
    if (lane == 0) {  // ---- producer: basis rows are older than the predecessor kernel, no wait ----
      int it = 0;
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const int t = reverse ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const uint32_t bytes = (uint32_t)len * sizeof(T);
        for (int gg = 0; gg < ngroups; ++gg, ++it) {
          const int g = reverse ? ngroups - 1 - gg : gg;
          const int s = it % kStages;
          tma::mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
          const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
          tma::mbar_arrive_expect_tx(full + s, bytes * rows_here);
          T* dst = stages + (size_t)s * kGroup * TILE;
          for (int r = 0; r < rows_here; ++r)
            tma::bulk_g2s(dst + (size_t)r * TILE, src.row(g * kGroup + r) + tc0 * (long long)sizeof(T), bytes,
                          full + s);
        }
      }
    }
  } else {  // ---- consumers: warp w takes row g*8 + w of every stage ----
    tma::griddep_wait();  // x (and everything else in global memory) may come from the predecessor
    int it = 0;
    for (int tt = 0; tt < cr.ntiles; ++tt) {
      const int t = reverse ? cr.ntiles - 1 - tt : tt;
      const long long tc0 = cr.c0 + (long long)t * TILE;
      const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
      V xr[XV];
#pragma unroll
      for (int u = 0; u < XV; ++u) {
        const int e = (lane + 32 * u) * VN;
        const long long col = tc0 + e;
        T tmp[VN];
#pragma unroll
        for (int k = 0; k < VN; ++k) tmp[k] = T(0);
        if (e < len) {
          if (col + VN <= n) {
            vec_unpack(__ldg(reinterpret_cast<const V*>(x + col)), tmp);
          } else {
#pragma unroll
            for (int k = 0; k < VN; ++k)
              if (col + k < n) tmp[k] = x[col + k];
          }
        }
        xr[u] = vec_pack(tmp);
      }
      for (int gg = 0; gg < ngroups; ++gg, ++it) {
        const int g = reverse ? ngroups - 1 - gg : gg;
        const int s = it % kStages;
        tma::mbar_wait(full + s, (it / kStages) & 1);
        const int j = g * kGroup + warp;
        if (j < nrows) {
          const V* row = reinterpret_cast<const V*>(stages + ((size_t)s * kGroup + warp) * TILE);
          T a0 = T(0), a1 = T(0);
#pragma unroll
          for (int u = 0; u < XV; ++u) {
            if ((lane + 32 * u) * VN < len) {
              T q[VN], xx[VN];
              vec_unpack(row[lane + 32 * u], q);
              vec_unpack(xr[u], xx);
#pragma unroll
              for (int k = 0; k < VN; ++k) {
                if (u & 1)
                  a1 = fma(q[k], xx[k], a1);
                else
                  a0 = fma(q[k], xx[k], a0);
              }
            }
          }
          double sacc = warp_sum(static_cast<double>(a0) + static_cast<double>(a1));
          if (lane == 0) acc_s[j] += sacc;  // row j is always handled by this warp: no race
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(empty + s);
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nrows; j += blockDim.x) partials[(size_t)j * gridDim.x + blockIdx.x] = acc_s[j];
  if (!last_block_done(counter)) return;
  reduce_partials_and_epilogue<T>(nrows, partials, epi);
}

// ---------------------------------------------------------------------------------------------
struct CombineTmaArgs {
  long long n = 0;
  void* out = nullptr;
  void* out2 = nullptr;
  int nvec = 0;
  VecTerm vec[kMaxVecTerms];
  RowSource src;
  const double* coef0 = nullptr;  // coefficients of block 0 rows (already offset)
  double sign0 = 1.0;
  const double* coef1 = nullptr;
  double sign1 = 1.0;
  const double* out_div_ptr = nullptr;
  const double* out_mul_ptr = nullptr;
  double* partials = nullptr;
  unsigned int* counter = nullptr;
  int reverse = 0;  // walk tiles and row groups backwards (L2 "snake" order, see krylov.cu)
  Epi epi;
};

template <typename T, bool NORM>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_combine_tma(CombineTmaArgs a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int TILE = kConsumerThreads * VN;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stages = reinterpret_cast<T*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kGroup * TILE * sizeof(T));
  uint64_t* empty = full + kStages;
  T* coef = reinterpret_cast<T*>(empty + kStages);  // [nrows]
  __shared__ double red_smem[32];
  __shared__ T vcoef[kMaxVecTerms];
  __shared__ T oscale[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = a.src.n0 + a.src.n1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, kConsumerWarps);
    }
    tma::fence_barrier_init();
  }
  __syncthreads();
  tma::griddep_launch_dependents();

  const ColumnRange cr = block_columns<T>(a.n, TILE);
  const int ngroups = (nrows + kGroup - 1) / kGroup;
  const int reverse = a.reverse;
  double ss = 0.0;

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      int it = 0;
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const int t = reverse ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const uint32_t bytes = (uint32_t)len * sizeof(T);
        for (int gg = 0; gg < ngroups; ++gg, ++it) {
          const int g = reverse ? ngroups - 1 - gg : gg;
          const int s = it % kStages;
          tma::mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
          const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
          tma::mbar_arrive_expect_tx(full + s, bytes * rows_here);
          T* dst = stages + (size_t)s * kGroup * TILE;
          for (int r = 0; r < rows_here; ++r)
            tma::bulk_g2s(dst + (size_t)r * TILE, a.src.row(g * kGroup + r) + tc0 * (long long)sizeof(T), bytes,
                          full + s);
        }
      }
    }
  } else {
    const int tid = threadIdx.x;  // 0..255: one 16-byte vector of every staged row
    // coefficients and scalars come from the predecessor's epilogue: wait for it first
    tma::griddep_wait();
    if (tid == 0) {
      oscale[0] = a.out_mul_ptr ? static_cast<T>(*a.out_mul_ptr) : T(1);
      oscale[1] = a.out_div_ptr ? static_cast<T>(*a.out_div_ptr) : T(1);
    }
    for (int j = tid; j < nrows; j += kConsumerThreads)
      coef[j] = j < a.src.n0 ? static_cast<T>(a.sign0 * a.coef0[j]) : static_cast<T>(a.sign1 * a.coef1[j - a.src.n0]);
    if (tid < a.nvec) {
      const VecTerm& v = a.vec[tid];
      vcoef[tid] = static_cast<T>(v.coef_imm * (v.coef_ptr ? *v.coef_ptr : 1.0));
    }
    tma::named_bar_sync(2, kConsumerThreads);
    int it = 0;
    for (int tt = 0; tt < cr.ntiles; ++tt) {
      const int t = reverse ? cr.ntiles - 1 - tt : tt;
      const long long tc0 = cr.c0 + (long long)t * TILE;
      const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
      const long long col = tc0 + (long long)tid * VN;
      const bool live = tid * VN < len;
      const bool fullvec = live && col + VN <= a.n;
      T acc[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = T(0);
      if (live) {
        for (int v = 0; v < a.nvec; ++v) {
          T e[VN];
          if (fullvec) {
            vec_unpack(reinterpret_cast<const V*>(a.vec[v].ptr)[col / VN], e);
          } else {
#pragma unroll
            for (int k = 0; k < VN; ++k) e[k] = col + k < a.n ? static_cast<const T*>(a.vec[v].ptr)[col + k] : T(0);
          }
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[k] = fma(vcoef[v], e[k], acc[k]);
        }
      }
      for (int gg = 0; gg < ngroups; ++gg, ++it) {
        const int g = reverse ? ngroups - 1 - gg : gg;
        const int s = it % kStages;
        tma::mbar_wait(full + s, (it / kStages) & 1);
        const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
        if (live) {
          const V* st = reinterpret_cast<const V*>(stages + (size_t)s * kGroup * TILE) + tid;
          const T* cf = coef + g * kGroup;
          if (rows_here == kGroup) {
            V q[kGroup];
#pragma unroll
            for (int r = 0; r < kGroup; ++r) q[r] = st[(size_t)r * (TILE / VN)];
#pragma unroll
            for (int r = 0; r < kGroup; ++r) {
              T e[VN];
              vec_unpack(q[r], e);
#pragma unroll
              for (int k = 0; k < VN; ++k) acc[k] = fma(cf[r], e[k], acc[k]);
            }
          } else {
            for (int r = 0; r < rows_here; ++r) {
              T e[VN];
              vec_unpack(st[(size_t)r * (TILE / VN)], e);
#pragma unroll
              for (int k = 0; k < VN; ++k) acc[k] = fma(cf[r], e[k], acc[k]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(empty + s);
      }
      if (live) {
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          acc[k] = acc[k] * oscale[0] / oscale[1];
          if (NORM && col + k < a.n) ss += static_cast<double>(acc[k] * acc[k]);
        }
        if (fullvec) {
          reinterpret_cast<V*>(a.out)[col / VN] = vec_pack(acc);
          if (a.out2) reinterpret_cast<V*>(a.out2)[col / VN] = vec_pack(acc);
        } else {
#pragma unroll
          for (int k = 0; k < VN; ++k)
            if (col + k < a.n) {
              static_cast<T*>(a.out)[col + k] = acc[k];
              if (a.out2) static_cast<T*>(a.out2)[col + k] = acc[k];
            }
        }
      }
    }
  }
  if (!NORM) return;
  double bs = block_sum(ss, red_smem);
  if (threadIdx.x == 0) a.partials[blockIdx.x] = bs;
  if (!last_block_done(a.counter)) return;
  double s = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(a.partials + b);
  s = block_sum(s, red_smem);
  if (threadIdx.x == 0) a.epi.red[0] = s;
  __syncthreads();
  run_epilogue<T>(a.epi);
}

// EPT consecutive elements of T as one 4/8/16-byte shared-memory load (conflict-free).
template <typename T, int EPT>
__device__ __forceinline__ void load_ept(const T* p, T (&e)[EPT]) {
  constexpr int BYTES = EPT * (int)sizeof(T);
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "EPT * sizeof(T) must be 4, 8 or 16");
  if constexpr (BYTES == 4) {
    e[0] = p[0];
  } else if constexpr (BYTES == 8) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    memcpy(e, &v, 8);
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    memcpy(e, &v, 16);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused second Gram-Schmidt pass: x' = x - sum_j c_j row_j (written to `out`), then
// red[j] = <row_j, x'> from the tile that is still resident in shared memory.
// EPT = elements per consumer thread in sweep 1; TILE = 256 * EPT columns.
template <typename T, int EPT>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_project_tma(RowSource src, int nrows, const T* x, T* out, long long n, const double* __restrict__ coef_in,
              double sign, double* __restrict__ partials, unsigned int* counter, Epi epi, int reverse) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int TILE = kConsumerThreads * EPT;
  constexpr int LV = TILE / (32 * VN);  // 16-byte vectors per lane in sweep 2
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ngroups = (nrows + kGroup - 1) / kGroup;
  T* tile_s = reinterpret_cast<T*>(smem_raw);                    // [ngroups*8][TILE]
  T* xs = tile_s + (size_t)ngroups * kGroup * TILE;              // [TILE]   x' of the tile
  T* coef = xs + TILE;                                           // [ngroups*8]
  uint64_t* full = reinterpret_cast<uint64_t*>(
      smem_raw + (((size_t)ngroups * kGroup * (TILE + 1) + TILE) * sizeof(T) + 15) / 16 * 16);  // [ngroups]
  uint64_t* tile_free = full + ngroups;
  double* acc_s = reinterpret_cast<double*>(tile_free + 1);      // [nrows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int g = 0; g < ngroups; ++g) tma::mbar_init(full + g, 1);
    tma::mbar_init(tile_free, kConsumerWarps);
    tma::fence_barrier_init();
  }
  for (int j = threadIdx.x; j < nrows; j += blockDim.x) acc_s[j] = 0.0;
  __syncthreads();
  tma::griddep_launch_dependents();

  const ColumnRange cr = block_columns<T>(n, TILE);

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const int t = reverse ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const uint32_t bytes = (uint32_t)len * sizeof(T);
        tma::mbar_wait(tile_free, (tt & 1) ^ 1);  // both sweeps of the previous tile are done
        for (int gg = 0; gg < ngroups; ++gg) {
          const int g = reverse ? ngroups - 1 - gg : gg;
          const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
          tma::mbar_arrive_expect_tx(full + g, bytes * rows_here);
          for (int r = 0; r < rows_here; ++r)
            tma::bulk_g2s(tile_s + (size_t)(g * kGroup + r) * TILE,
                          src.row(g * kGroup + r) + tc0 * (long long)sizeof(T), bytes, full + g);
        }
      }
    }
  } else {
    const int tid = threadIdx.x;
    tma::griddep_wait();  // coefficients come from the predecessor's epilogue
    for (int j = tid; j < ngroups * kGroup; j += kConsumerThreads)
      coef[j] = j < nrows ? static_cast<T>(sign * coef_in[j]) : T(0);
    tma::named_bar_sync(2, kConsumerThreads);
    for (int tt = 0; tt < cr.ntiles; ++tt) {
      const int t = reverse ? cr.ntiles - 1 - tt : tt;
      const long long tc0 = cr.c0 + (long long)t * TILE;
      const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
      // ---- sweep 1: x' = x + sum_j coef_j row_j (thread <-> EPT consecutive columns) ----
      T acc[EPT];
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const long long col = tc0 + (long long)tid * EPT + k;
        acc[k] = (tid * EPT + k < len && col < n) ? x[col] : T(0);
      }
      for (int gg = 0; gg < ngroups; ++gg) {
        const int g = reverse ? ngroups - 1 - gg : gg;
        tma::mbar_wait(full + g, tt & 1);
        const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
        const T* st = tile_s + (size_t)g * kGroup * TILE + tid * EPT;
        const T* cf = coef + g * kGroup;
        if (tid * EPT < len) {
#pragma unroll
          for (int r = 0; r < kGroup; ++r) {
            if (r < rows_here) {
              T e[EPT];
              load_ept<T, EPT>(st + (size_t)r * TILE, e);
#pragma unroll
              for (int k = 0; k < EPT; ++k) acc[k] = fma(cf[r], e[k], acc[k]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const long long col = tc0 + (long long)tid * EPT + k;
        const bool ok = tid * EPT + k < len && col < n;
        xs[tid * EPT + k] = ok ? acc[k] : T(0);
        if (ok) out[col] = acc[k];
      }
      tma::named_bar_sync(1, kConsumerThreads);
      // ---- sweep 2: red[j] += <row_j, x'>, warp w owns rows w, w+8, ... ----
      V xv[LV];
#pragma unroll
      for (int u = 0; u < LV; ++u) xv[u] = reinterpret_cast<const V*>(xs)[lane + 32 * u];
      // four rows per iteration: four independent reduction chains hide the shuffle latency
      for (int j = warp; j < nrows; j += 4 * kConsumerWarps) {
        T a[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int jj = j + q4 * kConsumerWarps;
          a[q4] = T(0);
          if (jj < nrows) {
            const V* row = reinterpret_cast<const V*>(tile_s + (size_t)jj * TILE);
#pragma unroll
            for (int u = 0; u < LV; ++u) {
              if ((lane + 32 * u) * VN < len) {
                T q[VN], xx[VN];
                vec_unpack(row[lane + 32 * u], q);
                vec_unpack(xv[u], xx);
#pragma unroll
                for (int k = 0; k < VN; ++k) a[q4] = fma(q[k], xx[k], a[q4]);
              }
            }
          }
        }
        double d[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) d[q4] = static_cast<double>(a[q4]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) d[q4] += __shfl_xor_sync(0xffffffffu, d[q4], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (j + q4 * kConsumerWarps < nrows) acc_s[j + q4 * kConsumerWarps] += d[q4];
        }
      }
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(tile_free);
      tma::named_bar_sync(1, kConsumerThreads);  // xs is rewritten by the next tile's sweep 1
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nrows; j += blockDim.x) partials[(size_t)j * gridDim.x + blockIdx.x] = acc_s[j];
  if (!last_block_done(counter)) return;
  reduce_partials_and_epilogue<T>(nrows, partials, epi);
}

}  // namespace bl
