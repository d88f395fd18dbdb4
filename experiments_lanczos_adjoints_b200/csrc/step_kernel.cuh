// One Krylov step's worth of basis streaming in ONE launch (sm_100a): the symmetric loops of the forward
// (arnoldi.py:87-98) and of the adjoint (arnoldi.py:201-219) are, per step and after the operator call,
//
//   phase 0   red0[j] = <few_j, x0>                     two or three neighbouring basis rows      (k_dots_few)
//   phase 1   out1 = (sum_k a_k vec_k) / div ; red1[j] = <row_j, out1>  over the active rows      (k_xdots_tma)
//   phase 2   out2 = out1 + sum_j c_j row_j  (+ ||out2||^2)             over the active rows      (k_combine_tma)
//
// with a grid-wide reduction between the phases whose result (Gram-Schmidt coefficients) the next phase
// needs.  As three launches each reduction costs a kernel boundary: ramp, last-block pass, drain, launch
// (8-10 us against 2-60 us of streaming).  Here the three phases run in one cooperative grid: blocks meet at
// a counter in L2, every block (few values) or block j (value j, many values) reduces the per-block partials
// in a fixed order, and EVERY block runs the epilogue on identical numbers (the global writes are redundant
// and identical).  The TMA producer thread does not take part in the barriers: while the consumers wait it
// already fills the 96 KB ring with the next phase's rows, so HBM stays busy across the reductions.
//
// Each block owns the same contiguous column range in all phases; thread t of the 256 consumers owns column
// vector t of every tile, so phase 2 reads back what the same thread wrote in phase 1 (program order, no
// fence), and phase 2 walks the tiles in the opposite direction (it starts on the rows phase 1 read last).
#pragma once

#include "stream_kernels.cuh"

namespace bl {

constexpr int kFewMax = 4;

struct StepArgs {
  long long n = 0;
  // ---- phase 0 ----
  int few_n = 0;
  const void* few_row[kFewMax] = {nullptr, nullptr, nullptr, nullptr};
  const void* few_x = nullptr;
  Epi epi0;
  // ---- phase 1 ----
  RowSource src1;
  int nrows1 = 0;
  void* out1 = nullptr;
  int nvec = 0;
  VecTerm vec[kXTerms];
  const double* out_div_ptr = nullptr;
  Epi epi1;
  // ---- phase 2 ----
  RowSource src2;
  int nrows2 = 0;
  const double* coef2 = nullptr;  // coefficient of row j: sign2 * coef2[j]
  double sign2 = 1.0;
  void* out2 = nullptr;
  int norm = 0;  // ||out2||^2 -> epi2 (run by the last block to leave)
  Epi epi2;
  // ---- reductions ----
  double* partials = nullptr;       // [rows][gridDim.x]
  double* red_g = nullptr;          // [rows] reduced values of the distributed path
  double* partials_norm = nullptr;  // [gridDim.x]
  unsigned int* bar = nullptr;      // arrival counter of the in-kernel barriers (a multiple of gridDim.x between launches)
  unsigned int* exit_counter = nullptr;
  int reverse = 0;       // direction of phase 1; phase 2 walks the other way
  int wait_row = 1 << 30;  // rows >= wait_row of src1 are the predecessor kernel's output: the producer
                           // executes griddepcontrol.wait before it copies them
};

namespace step {

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Barrier among the consumer threads of all blocks (the producer warp is not involved).  The counter only
// grows inside a launch: block tickets of barrier k lie in [k G, (k+1) G), so the release value is the next
// multiple of G above the own ticket -- no generation flag, no reset, no host-side epoch.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, int tid) {
  __threadfence();
  tma::named_bar_sync(1, kConsumerThreads);
  if (tid == 0) {
    const unsigned int G = gridDim.x;
    const unsigned int ticket = atomicAdd(bar, 1u);
    const unsigned int target = (ticket / G + 1u) * G;
    while ((int)(ld_acquire(bar) - target) < 0) {
    }
    __threadfence();
  }
  tma::named_bar_sync(1, kConsumerThreads);
}

// fixed-order sum of one row of per-block partials by one warp; all loads of a lane are issued before the adds
__device__ __forceinline__ double row_sum(const double* __restrict__ p, int G, int lane) {
  double s = 0.0;
  if ((G & 1) == 0 && G <= 320) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const int nv = G >> 1;
    double2 v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int i = lane + 32 * k;
      v[k] = __ldcg(p2 + (i < nv ? i : 0));
      if (i >= nv) v[k] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) s += v[k].x + v[k].y;
  } else {
    for (int k = lane; k < G; k += 32) s += __ldcg(p + k);
  }
  return warp_sum(s);
}

// Sum over blocks of `nrows` per-block values (src: shared memory of this block) -> red_s[0..nrows) in every
// block, bit-identical everywhere (fixed order).  Few values: one barrier, every block reduces all of them.
// Many: block j reduces value j (rows j, j+G, ...), a second barrier, everybody reads the results.
__device__ __forceinline__ void grid_reduce(const StepArgs& a, int nrows, const double* src, double* red_s, int tid) {
  const int G = gridDim.x, b = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < nrows; j += kConsumerThreads) __stcg(a.partials + (size_t)j * G + b, src[j]);
  grid_barrier(a.bar, tid);
  if (nrows * G <= 2560) {
    for (int j = warp; j < nrows; j += kConsumerWarps) {
      const double s = row_sum(a.partials + (size_t)j * G, G, lane);
      if (lane == 0) red_s[j] = s;
    }
  } else {
    for (int j = b + warp * G; j < nrows; j += kConsumerWarps * G) {
      const double s = row_sum(a.partials + (size_t)j * G, G, lane);
      if (lane == 0) __stcg(a.red_g + j, s);
    }
    grid_barrier(a.bar, tid);
    for (int j = tid; j < nrows; j += kConsumerThreads) red_s[j] = __ldcg(a.red_g + j);
  }
  tma::named_bar_sync(1, kConsumerThreads);
}

struct ConsumerSync {
  __device__ __forceinline__ void operator()() const { tma::named_bar_sync(1, kConsumerThreads); }
};

}  // namespace step

template <typename T, int TILE>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_step_tma(const __grid_constant__ StepArgs a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int XV = TILE / (32 * VN);  // x vectors per lane
  static_assert(TILE == kConsumerThreads * VN, "one 16-byte column vector per consumer thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stages = reinterpret_cast<T*>(smem_raw);                       // [kStages][kGroup][TILE]
  T* xs = stages + (size_t)kStages * kGroup * TILE;                 // [2][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(xs + 2 * TILE);
  uint64_t* empty = full + kStages;
  double* acc_s = reinterpret_cast<double*>(empty + kStages + 2);   // [max(nrows1, kFewMax * 8)]
  const int nacc = a.nrows1 > kFewMax * kConsumerWarps ? a.nrows1 : kFewMax * kConsumerWarps;
  double* red_s = acc_s + nacc;                                     // [max(nrows1, kFewMax)]
  T* coef_s = reinterpret_cast<T*>(red_s + (a.nrows1 > kFewMax ? a.nrows1 : kFewMax));  // [ceil8(nrows2)]
  __shared__ double red_smem[32];

  const int nrows1 = a.nrows1, nrows2 = a.nrows2;
  const int dir1 = a.reverse, dir2 = a.reverse ^ 1;
  const long long n = a.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, kConsumerWarps);
    }
    tma::fence_barrier_init();
  }
  for (int j = threadIdx.x; j < nacc; j += blockDim.x) acc_s[j] = 0.0;
  __syncthreads();
  tma::griddep_launch_dependents();

  const ColumnRange cr = block_columns<T>(n, TILE);
  const int ngroups1 = (nrows1 + kGroup - 1) / kGroup;
  const int ngroups2 = (nrows2 + kGroup - 1) / kGroup;
  double ss = 0.0;

  if (warp == kConsumerWarps) {
    if (lane == 0) {  // ---- producer: the rows of phase 1, then the rows of phase 2, one ring ----
      bool waited = false;
      int it = 0;
      for (int phase = 1; phase <= 2; ++phase) {
        const RowSource& src = phase == 1 ? a.src1 : a.src2;
        const int nrows = phase == 1 ? nrows1 : nrows2;
        const int ngroups = phase == 1 ? ngroups1 : ngroups2;
        const int dir = phase == 1 ? dir1 : dir2;
        for (int tt = 0; tt < cr.ntiles; ++tt) {
          const int t = dir ? cr.ntiles - 1 - tt : tt;
          const long long tc0 = cr.c0 + (long long)t * TILE;
          const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
          const uint32_t bytes = (uint32_t)len * sizeof(T);
          for (int gg = 0; gg < ngroups; ++gg, ++it) {
            const int g = dir ? ngroups - 1 - gg : gg;
            const int rows_here = nrows - g * kGroup < kGroup ? nrows - g * kGroup : kGroup;
            if (!waited && (phase == 2 || g * kGroup + rows_here > a.wait_row)) {
              tma::griddep_wait();
              waited = true;
            }
            const int s = it % kStages;
            tma::mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
            tma::mbar_arrive_expect_tx(full + s, bytes * rows_here);
            T* dst = stages + (size_t)s * kGroup * TILE;
            for (int r = 0; r < rows_here; ++r)
              tma::bulk_g2s(dst + (size_t)r * TILE, src.row(g * kGroup + r) + tc0 * (long long)sizeof(T), bytes, full + s);
          }
        }
      }
    }
  } else {
    const int tid = threadIdx.x;  // 0..255: column vector tid of every tile
    tma::griddep_wait();          // everything below reads the predecessor's output
    step::ConsumerSync csync;

    // ================= phase 0: dots of a few rows with x0 =================
    if (a.few_n > 0) {
      T facc[kFewMax];
#pragma unroll
      for (int r = 0; r < kFewMax; ++r) facc[r] = T(0);
      const T* x0 = static_cast<const T*>(a.few_x);
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const long long c = cr.c0 + (long long)tt * TILE + (long long)tid * VN;
        if (c >= cr.c1 || c >= n) continue;
        T xx[VN];
        if (c + VN <= n) {
          vec_unpack(*reinterpret_cast<const V*>(x0 + c), xx);
        } else {
#pragma unroll
          for (int k = 0; k < VN; ++k) xx[k] = c + k < n ? x0[c + k] : T(0);
        }
#pragma unroll
        for (int r = 0; r < kFewMax; ++r) {
          if (r < a.few_n) {
            T q[VN];  // basis rows are zero-padded up to ld: the straddling vector is readable
            vec_unpack(*reinterpret_cast<const V*>(static_cast<const T*>(a.few_row[r]) + c), q);
#pragma unroll
            for (int k = 0; k < VN; ++k) facc[r] = fma(q[k], xx[k], facc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kFewMax; ++r) {
        if (r < a.few_n) {
          const double w = warp_sum(static_cast<double>(facc[r]));
          if (lane == 0) acc_s[r * kConsumerWarps + warp] = w;
        }
      }
      csync();
      if (tid < a.few_n) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) s += acc_s[tid * kConsumerWarps + w];
        red_s[tid] = s;  // this block's value of row tid
      }
      csync();
      if (tid < kFewMax * kConsumerWarps) acc_s[tid] = 0.0;  // phase 1 accumulates into acc_s
      step::grid_reduce(a, a.few_n, red_s, red_s, tid);
      {
        Epi e = a.epi0;
        e.red = red_s;
        run_epilogue_impl<T>(e, tid, kConsumerThreads, csync);
      }
      __threadfence();  // the coefficients go through global memory (every block writes the same values)
      csync();
    }

    // ================= phase 1: out1 = (sum of terms) / div, red1[j] = <row_j, out1> =================
    {
      T cv[kXTerms];
#pragma unroll
      for (int v = 0; v < kXTerms; ++v)
        cv[v] = v < a.nvec ? static_cast<T>(a.vec[v].coef_imm * (a.vec[v].coef_ptr ? __ldcg(a.vec[v].coef_ptr) : 1.0)) : T(0);
      const T oscale = a.out_div_ptr ? static_cast<T>(__ldcg(a.out_div_ptr)) : T(1);
      V term[kXTerms];
      auto load_terms = [&](int tt) {  // this thread's vector of every term, tile tt (columns past n read as zero)
        const int t = dir1 ? cr.ntiles - 1 - tt : tt;
        const long long c = cr.c0 + (long long)t * TILE + (long long)tid * VN;
        const bool inside = tt < cr.ntiles && c < cr.c1 && c + VN <= n;
#pragma unroll
        for (int v = 0; v < kXTerms; ++v) {
          T z[VN];
#pragma unroll
          for (int k = 0; k < VN; ++k) z[k] = T(0);
          if (v < a.nvec) {
            const T* p = static_cast<const T*>(a.vec[v].ptr) + c;
            if (inside) {
              term[v] = *reinterpret_cast<const V*>(p);
              continue;
            }
            if (tt < cr.ntiles && c < cr.c1)  // the vector that straddles n
#pragma unroll
              for (int k = 0; k < VN; ++k)
                if (c + k < n) z[k] = p[k];
          }
          term[v] = vec_pack(z);
        }
      };
      load_terms(0);
      int it = 0;
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const int t = dir1 ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const int b = tt & 1;
        {  // ---- build this thread's vector of the x tile ----
          T acc[VN];
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[k] = T(0);
#pragma unroll
          for (int v = 0; v < kXTerms; ++v) {
            if (v < a.nvec) {
              T e[VN];
              vec_unpack(term[v], e);
#pragma unroll
              for (int k = 0; k < VN; ++k) acc[k] = fma(cv[v], e[k], acc[k]);
            }
          }
          const long long c = tc0 + (long long)tid * VN;
          bool all_ok = true;
#pragma unroll
          for (int k = 0; k < VN; ++k) {
            const bool ok = tid * VN + k < len && c + k < n;
            acc[k] = ok ? acc[k] / oscale : T(0);
            all_ok = all_ok && ok;
          }
          reinterpret_cast<V*>(xs + (size_t)b * TILE)[tid] = vec_pack(acc);
          if (all_ok) {
            *reinterpret_cast<V*>(static_cast<T*>(a.out1) + c) = vec_pack(acc);
          } else {
#pragma unroll
            for (int k = 0; k < VN; ++k)
              if (tid * VN + k < len && c + k < n) static_cast<T*>(a.out1)[c + k] = acc[k];
          }
        }
        load_terms(tt + 1);  // in flight while this tile's rows are consumed (in place: own columns only)
        csync();
        V xr[XV];
#pragma unroll
        for (int u = 0; u < XV; ++u) xr[u] = reinterpret_cast<const V*>(xs + (size_t)b * TILE)[lane + 32 * u];
        for (int gg = 0; gg < ngroups1; ++gg, ++it) {
          const int g = dir1 ? ngroups1 - 1 - gg : gg;
          const int s = it % kStages;
          tma::mbar_wait(full + s, (it / kStages) & 1);
          const int j = g * kGroup + warp;
          if (j < nrows1) {
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
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(empty + s);
            double sacc = warp_sum(static_cast<double>(a0) + static_cast<double>(a1));
            if (lane == 0) acc_s[j] += sacc;  // row j is always handled by this warp: no race
          } else {
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(empty + s);
          }
        }
      }
      csync();
      step::grid_reduce(a, nrows1, acc_s, red_s, tid);
      {
        Epi e = a.epi1;
        e.red = red_s;
        run_epilogue_impl<T>(e, tid, kConsumerThreads, csync);
      }
      __threadfence();
      csync();

      // ================= phase 2: out2 = out1 + sum_j c_j row_j (+ ||out2||^2) =================
      for (int j = tid; j < ngroups2 * kGroup; j += kConsumerThreads)
        coef_s[j] = j < nrows2 ? static_cast<T>(a.sign2 * __ldcg(a.coef2 + j)) : T(0);
      csync();
      auto load_x = [&](int tt) -> V {  // out1 as this very thread wrote it in phase 1
        T z[VN];
#pragma unroll
        for (int k = 0; k < VN; ++k) z[k] = T(0);
        if (tt < cr.ntiles) {
          const int t = dir2 ? cr.ntiles - 1 - tt : tt;
          const long long c = cr.c0 + (long long)t * TILE + (long long)tid * VN;
          if (c < cr.c1 && c + VN <= n) return *reinterpret_cast<const V*>(static_cast<const T*>(a.out1) + c);
          if (c < cr.c1)
#pragma unroll
            for (int k = 0; k < VN; ++k)
              if (c + k < n) z[k] = static_cast<const T*>(a.out1)[c + k];
        }
        return vec_pack(z);
      };
      V xnext = load_x(0);
      for (int tt = 0; tt < cr.ntiles; ++tt) {
        const int t = dir2 ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const long long col = tc0 + (long long)tid * VN;
        const bool live = tid * VN < len;
        const bool fullvec = live && col + VN <= n;
        T acc[VN];
        vec_unpack(xnext, acc);
        xnext = load_x(tt + 1);
        for (int gg = 0; gg < ngroups2; ++gg, ++it) {
          const int g = dir2 ? ngroups2 - 1 - gg : gg;
          const int s = it % kStages;
          tma::mbar_wait(full + s, (it / kStages) & 1);
          const int rows_here = nrows2 - g * kGroup < kGroup ? nrows2 - g * kGroup : kGroup;
          if (live) {
            const V* st = reinterpret_cast<const V*>(stages + (size_t)s * kGroup * TILE) + tid;
            const T* cf = coef_s + g * kGroup;
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
          if (a.norm) {
#pragma unroll
            for (int k = 0; k < VN; ++k)
              if (col + k < n) ss += static_cast<double>(acc[k] * acc[k]);
          }
          if (fullvec) {
            *reinterpret_cast<V*>(static_cast<T*>(a.out2) + col) = vec_pack(acc);
          } else {
#pragma unroll
            for (int k = 0; k < VN; ++k)
              if (col + k < n) static_cast<T*>(a.out2)[col + k] = acc[k];
          }
        }
      }
    }
  }
  // ---- exit: the last block to leave re-arms the barrier counter and (norm) closes the reduction ----
  const double bs = a.norm ? block_sum(ss, red_smem) : 0.0;
  if (a.norm && threadIdx.x == 0) a.partials_norm[blockIdx.x] = bs;
  if (!last_block_done(a.exit_counter)) return;
  if (threadIdx.x == 0) *a.bar = 0u;
  if (!a.norm) return;
  double s = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(a.partials_norm + b);
  s = block_sum(s, red_smem);
  if (threadIdx.x == 0) a.epi2.red[0] = s;
  __syncthreads();
  run_epilogue<T>(a.epi2);
}

}  // namespace bl
