// One Krylov step's worth of basis streaming in ONE launch, for up to kStepBatch independent runs (sm_100a).
//
// Per step and after the operator call, the symmetric loops of the forward (arnoldi.py:87-98) and of the adjoint
// (arnoldi.py:201-219) are
//
//   phase 0   red0[j] = <few_j, x0>                     two or three neighbouring basis rows      (k_dots_few)
//   phase 1   out1 = (sum_k a_k vec_k) / div ; red1[j] = <row_j, out1>  over the active rows      (k_xdots_tma)
//   phase 2   out2 = out1 + sum_j c_j row_j  (+ ||out2||^2)             over the active rows      (k_combine_tma)
//
// with a grid-wide reduction between the phases whose result (Gram-Schmidt coefficients) the next phase needs.
// As separate launches every reduction costs a kernel boundary (ramp, last-block pass, drain, launch: 8 us
// against 2-60 us of streaming), and a run is a chain of 800 of them.  Here the phases run in one cooperative grid:
// blocks meet at a counter in L2, every block (few values) or block j (value j, many values) reduces the per-block
// partials in a fixed order, and EVERY block runs the epilogue on identical numbers (its global writes are
// redundant and identical).  The TMA producer thread takes no part in the barriers: while the consumers wait it
// fills the 96 KB ring with the next phase's rows.
//
// Measured (block 0's time stamps, bl_step_trace_*), an in-kernel reduction costs ~6 us -- no less than a kernel
// boundary with programmatic dependent launch.  What pays is the BATCH: the lockstep drivers
// (bl_arnoldi_{forward,adjoint}_batch: Hutchinson probes, initial conditions) hand the step of up to kStepBatch runs
// to one launch; each phase loops over the runs, and ONE barrier per phase serves all of them, so a run's share
// of every fixed cost is divided by the batch size.
//
// Each block owns the same contiguous column range in all phases; thread t of the 256 consumers owns column
// vector t of every tile, so phase 2 reads back what the same thread wrote in phase 1 (program order, no
// fence), and phase 2 walks the tiles in the opposite direction (it starts on the rows phase 1 read last).
#pragma once

#include "stream_kernels.cuh"

namespace bl {

constexpr int kFewMax = 4;
constexpr int kStepBatch = 4;
constexpr int kFewSlots = kFewMax * kConsumerWarps + 8;  // per-run scratch of phase 0: warp partials, then block values

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
  // ---- reductions (per run) ----
  double* partials = nullptr;       // [rows][gridDim.x]
  double* red_g = nullptr;          // [rows] reduced values of the distributed path
  double* partials_norm = nullptr;  // [gridDim.x]
  int wait_row = 1 << 30;  // rows >= wait_row of src1 are the predecessor kernel's output: the producer
                           // executes griddepcontrol.wait before it copies them
};

struct StepBatch {
  int count = 1;
  int reverse = 0;                      // direction of phase 1; phase 2 walks the other way
  int trace_slot = -1;                  // >= 0: block 0 writes its time stamps to g_step_trace[trace_slot]
  int acc_stride = 0, coef_stride = 0;  // per-run shared-memory strides (doubles / elements)
  unsigned int* bar = nullptr;          // arrival counter of the in-kernel barriers (0 between launches)
  unsigned int* exit_counter = nullptr;
  StepArgs a[kStepBatch];
};

// Optional time stamps of block 0 (bl_step_trace_*): 8 stamps per launch -- globaltimer at entry, then SM clocks
// after the dependency wait, phase 0's loads, phase 0's reduction, phase 1's stream, phase 1's reduction, phase 2, exit.
constexpr int kTraceStamps = 8;
constexpr int kTraceLaunches = 2048;
__device__ unsigned long long g_step_trace[kTraceStamps * kTraceLaunches];

namespace step {

__device__ __forceinline__ void stamp(int slot, int k, int tid) {
  if (slot >= 0 && blockIdx.x == 0 && tid == 0) {
    unsigned long long t;
    if (k == 0)
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    else
      t = (unsigned long long)clock64();
    g_step_trace[(size_t)slot * kTraceStamps + k] = t;
  }
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct ConsumerSync {
  __device__ __forceinline__ void operator()() const { tma::named_bar_sync(1, kConsumerThreads); }
};

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

// Sum over blocks of the per-block values of every run: run p has nvals(p) values at vals(p)[0..), reduced in
// place -- bit-identical in every block (fixed order).  Few values in total: one barrier, every block reduces all of
// them.  Many: block j reduces value j (j, j+G, ... over the runs' values laid end to end), a second barrier,
// everybody reads the results.  PHASE selects which count of StepArgs is meant (0: few_n, 1: nrows1).
template <int PHASE>
__device__ __forceinline__ int nvals_of(const StepArgs& a) {
  return PHASE == 0 ? a.few_n : a.nrows1;
}

template <int PHASE>
__device__ __forceinline__ void grid_reduce(const StepBatch& B, double* acc_base, int val_off, int tid) {
  const int G = gridDim.x, b = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  int total = 0;
  for (int p = 0; p < B.count; ++p) {
    const int nv = nvals_of<PHASE>(B.a[p]);
    const double* src = acc_base + (size_t)p * B.acc_stride + val_off;
    for (int j = tid; j < nv; j += kConsumerThreads) __stcg(B.a[p].partials + (size_t)j * G + b, src[j]);
    total += nv;
  }
  grid_barrier(B.bar, tid);
  const bool everyone = total * G <= 2560;
  for (int v = everyone ? warp : b + warp * G; v < total; v += everyone ? kConsumerWarps : kConsumerWarps * G) {
    int p = 0, j = v;
    while (j >= nvals_of<PHASE>(B.a[p])) j -= nvals_of<PHASE>(B.a[p++]);
    const double s = row_sum(B.a[p].partials + (size_t)j * G, G, lane);
    if (lane == 0) {
      if (everyone)
        acc_base[(size_t)p * B.acc_stride + val_off + j] = s;
      else
        __stcg(B.a[p].red_g + j, s);
    }
  }
  if (!everyone) {
    grid_barrier(B.bar, tid);
    for (int p = 0; p < B.count; ++p) {
      double* dst = acc_base + (size_t)p * B.acc_stride + val_off;
      for (int j = tid; j < nvals_of<PHASE>(B.a[p]); j += kConsumerThreads) dst[j] = __ldcg(B.a[p].red_g + j);
    }
  }
  tma::named_bar_sync(1, kConsumerThreads);
}

}  // namespace step

template <typename T, int TILE>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_step_tma(const __grid_constant__ StepBatch B) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  constexpr int XV = TILE / (32 * VN);  // x vectors per lane
  static_assert(TILE == kConsumerThreads * VN, "one 16-byte column vector per consumer thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stages = reinterpret_cast<T*>(smem_raw);                       // [kStages][kGroup][TILE]
  T* xs = stages + (size_t)kStages * kGroup * TILE;                 // [2][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(xs + 2 * TILE);
  uint64_t* empty = full + kStages;
  double* acc_s = reinterpret_cast<double*>(empty + kStages + 2);   // [count][acc_stride]: dots of phase 1, in place reduced
  T* coef_s = reinterpret_cast<T*>(acc_s + (size_t)B.count * B.acc_stride);  // [count][coef_stride]
  __shared__ double red_smem[32];

  const int P = B.count;
  const int dir1 = B.reverse, dir2 = B.reverse ^ 1;
  const long long n = B.a[0].n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, kConsumerWarps);
    }
    tma::fence_barrier_init();
  }
  for (int j = threadIdx.x; j < P * B.acc_stride; j += blockDim.x) acc_s[j] = 0.0;
  __syncthreads();
  tma::griddep_launch_dependents();

  const ColumnRange cr = block_columns<T>(n, TILE);

  if (warp == kConsumerWarps) {
    if (lane == 0) {  // ---- producer: the rows of phase 1 of every run, then the rows of phase 2, one ring ----
      bool waited = false;
      int it = 0;
      for (int phase = 1; phase <= 2; ++phase) {
        const int dir = phase == 1 ? dir1 : dir2;
        for (int pp = 0; pp < P; ++pp) {
          const int p = phase == 1 ? pp : P - 1 - pp;  // phase 2 starts with the run whose rows were read last (L2)
          const StepArgs& a = B.a[p];
          const RowSource& src = phase == 1 ? a.src1 : a.src2;
          const int nrows = phase == 1 ? a.nrows1 : a.nrows2;
          const int ngroups = (nrows + kGroup - 1) / kGroup;
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
                tma::bulk_g2s(dst + (size_t)r * TILE, src.row(g * kGroup + r) + tc0 * (long long)sizeof(T), bytes,
                              full + s);
            }
          }
        }
      }
    }
  } else {
    const int tid = threadIdx.x;  // 0..255: column vector tid of every tile
    step::stamp(B.trace_slot, 0, tid);
    tma::griddep_wait();          // everything below reads the predecessor's output
    step::ConsumerSync csync;
    step::stamp(B.trace_slot, 1, tid);

    // ================= phase 0: dots of a few rows with x0, every run =================
    bool any_few = false;
    for (int p = 0; p < P; ++p) any_few = any_few || B.a[p].few_n > 0;
    if (any_few) {
      for (int p = 0; p < P; ++p) {
        const StepArgs& a = B.a[p];
        if (a.few_n <= 0) continue;
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
        double* scratch = acc_s + (size_t)p * B.acc_stride;
#pragma unroll
        for (int r = 0; r < kFewMax; ++r) {
          if (r < a.few_n) {
            const double w = warp_sum(static_cast<double>(facc[r]));
            if (lane == 0) scratch[r * kConsumerWarps + warp] = w;
          }
        }
      }
      csync();
      for (int p = 0; p < P; ++p) {
        double* scratch = acc_s + (size_t)p * B.acc_stride;
        if (tid < B.a[p].few_n) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < kConsumerWarps; ++w) s += scratch[tid * kConsumerWarps + w];
          scratch[kFewMax * kConsumerWarps + tid] = s;  // this block's value of row tid
        }
      }
      csync();
      step::stamp(B.trace_slot, 2, tid);
      step::grid_reduce<0>(B, acc_s, kFewMax * kConsumerWarps, tid);
      for (int p = 0; p < P; ++p) {
        if (B.a[p].few_n <= 0) continue;
        run_epilogue_impl<T>(B.a[p].epi0, acc_s + (size_t)p * B.acc_stride + kFewMax * kConsumerWarps, tid,
                             kConsumerThreads, csync);
      }
      __threadfence();  // the coefficients go through global memory (every block writes the same values)
      csync();
      for (int p = 0; p < P; ++p)
        if (tid < kFewSlots) acc_s[(size_t)p * B.acc_stride + tid] = 0.0;  // phase 1 accumulates into acc_s
      csync();
      step::stamp(B.trace_slot, 3, tid);
    }

    // ================= phase 1: out1 = (sum of terms) / div, red1[j] = <row_j, out1>, every run =================
    int it = 0, xt = 0;
    for (int p = 0; p < P; ++p) {
      const StepArgs& a = B.a[p];
      const int nrows1 = a.nrows1;
      const int ngroups1 = (nrows1 + kGroup - 1) / kGroup;
      double* acc_p = acc_s + (size_t)p * B.acc_stride;
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
            const T* ptr = static_cast<const T*>(a.vec[v].ptr) + c;
            if (inside) {
              term[v] = *reinterpret_cast<const V*>(ptr);
              continue;
            }
            if (tt < cr.ntiles && c < cr.c1)  // the vector that straddles n
#pragma unroll
              for (int k = 0; k < VN; ++k)
                if (c + k < n) z[k] = ptr[k];
          }
          term[v] = vec_pack(z);
        }
      };
      load_terms(0);
      for (int tt = 0; tt < cr.ntiles; ++tt, ++xt) {
        const int t = dir1 ? cr.ntiles - 1 - tt : tt;
        const long long tc0 = cr.c0 + (long long)t * TILE;
        const int len = (int)((cr.c1 - tc0) < TILE ? (cr.c1 - tc0) : TILE);
        const int b = xt & 1;
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
            if (lane == 0) acc_p[j] += sacc;  // row j is always handled by this warp: no race
          } else {
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(empty + s);
          }
        }
      }
    }
    csync();
    step::stamp(B.trace_slot, 4, tid);
    step::grid_reduce<1>(B, acc_s, 0, tid);
    for (int p = 0; p < P; ++p)
      run_epilogue_impl<T>(B.a[p].epi1, acc_s + (size_t)p * B.acc_stride, tid, kConsumerThreads, csync);
    __threadfence();
    csync();

    // ================= phase 2: out2 = out1 + sum_j c_j row_j (+ ||out2||^2), every run =================
    for (int p = 0; p < P; ++p) {
      const StepArgs& a = B.a[p];
      const int padded = (a.nrows2 + kGroup - 1) / kGroup * kGroup;
      T* cf = coef_s + (size_t)p * B.coef_stride;
      for (int j = tid; j < padded; j += kConsumerThreads)
        cf[j] = j < a.nrows2 ? static_cast<T>(a.sign2 * __ldcg(a.coef2 + j)) : T(0);
    }
    csync();
    step::stamp(B.trace_slot, 5, tid);
    for (int pp = 0; pp < P; ++pp) {
      const int p = P - 1 - pp;  // the producer's order
      const StepArgs& a = B.a[p];
      double ss = 0.0;
      const int nrows2 = a.nrows2;
      const int ngroups2 = (nrows2 + kGroup - 1) / kGroup;
      const T* coef_p = coef_s + (size_t)p * B.coef_stride;
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
            const T* cf = coef_p + g * kGroup;
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
      if (a.norm) {  // acc_s is free by now: warp sums of ||out2||^2, added up at the exit
        ss = warp_sum(ss);
        if (lane == 0) acc_s[(size_t)p * B.acc_stride + warp] = ss;
      }
    }
    step::stamp(B.trace_slot, 6, tid);
  }
  // ---- exit: the last block to leave re-arms the barrier counter and (norm) closes the reductions ----
  __syncthreads();
  if ((int)threadIdx.x < P && B.a[threadIdx.x].norm) {
    const double* w = acc_s + (size_t)threadIdx.x * B.acc_stride;
    double bs = 0.0;
#pragma unroll
    for (int k = 0; k < kConsumerWarps; ++k) bs += w[k];
    B.a[threadIdx.x].partials_norm[blockIdx.x] = bs;
  }
  step::stamp(B.trace_slot, 7, threadIdx.x);
  if (!last_block_done(B.exit_counter)) return;
  if (threadIdx.x == 0) *B.bar = 0u;
  for (int p = 0; p < P; ++p) {
    if (!B.a[p].norm) continue;
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(B.a[p].partials_norm + b);
    s = block_sum(s, red_smem);
    if (threadIdx.x == 0) B.a[p].epi2.red[0] = s;
    __syncthreads();
    run_epilogue<T>(B.a[p].epi2);
    __syncthreads();
  }
}

}  // namespace bl
