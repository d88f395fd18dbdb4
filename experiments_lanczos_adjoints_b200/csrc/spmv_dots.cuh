// The operator call of ONE run's Krylov step together with its neighbouring-row dots (sm_100a).
//
// One run alone takes its step from separate launches (an in-kernel grid barrier costs what a kernel boundary with
// programmatic dependent launch costs, step_kernel.cuh), and of those the dots of the operator's output with two or
// three basis rows -- <q_{i-1}, A q_i>, <q_i, A q_i> in the forward (arnoldi.py:87), <q_{idx-2..idx}, A^T lambda> in the
// adjoint (arnoldi.py:213) -- were the worst: k_dots_few moves 12-16 MB in 10-15 us, 200 launches and 13 % of a run.
// Here the SELL-32 SpMV computes them on the way: the lane that owns row r holds y[r] in a register, the basis
// entries q_j[r] are coalesced loads issued before the gather loop, and a block leaves ONE share per dot product in
// partials[j * gridDim.x + block] -- no fence, no atomic, no last-block pass in blocks that live a few microseconds
// (that variant was measured in round 1 and lost).  The NEXT kernel (k_xdots_tma, XDotsArgs::pre_*) adds the shares up in
// a fixed order in every block and runs the epilogue itself.  A PLAIN launch, like every operator kernel: the streaming
// kernels that follow prefetch basis rows before their dependency wait.
//
// y is bit-identical to k_sell_spmv_multi<T, 1, NORM, 6> (same chunks, same two accumulators).
#pragma once

#include "step_kernel.cuh"

namespace bl {

struct SpmvDotsArgs {
  const int64_t* slice_ptr = nullptr;
  const int32_t* col = nullptr;
  const void* val = nullptr;
  long long nslices = 0, nrows = 0, n_pad = 0;
  const void* x = nullptr;
  void* y = nullptr;
  const double* len = nullptr;  // NORM: q = x / *len (true division, arnoldi.py:80-81), y = A q
  void* q = nullptr;
  int few_n = 0;
  const void* few_row[kFewMax] = {nullptr, nullptr, nullptr, nullptr};
  int self = -1;  // NORM: few_row[self] IS q (written by this launch): its entry comes from the register
  double* partials = nullptr;  // [few_n][gridDim.x]
};

constexpr int kSpmvDotsMaxWarps = 32;

__device__ __forceinline__ int ld_stream_col(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <typename T, bool NORM, int W>
__global__ void __launch_bounds__(kSpmvDotsMaxWarps * 32)
k_sell_spmv_dots(const SpmvDotsArgs a) {
  __shared__ double share_s[kFewMax][kSpmvDotsMaxWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long slice = (long long)blockIdx.x * nwarps + warp;
  const long long r = slice * 32 + lane;
  const bool has_slice = slice < a.nslices;
  const bool live = has_slice && r < a.nrows;
  const T* x = static_cast<const T*>(a.x);
  T yv = T(0), qv[kFewMax];
#pragma unroll
  for (int j = 0; j < kFewMax; ++j) qv[j] = T(0);
  if (has_slice) {
    T inv = T(1), qself = T(0);
    if (NORM) {
      const T d = static_cast<T>(*a.len);
      inv = T(1) / d;
      qself = r < a.nrows ? x[r] * T(1) / d : T(0);
      if (r < a.n_pad) static_cast<T*>(a.q)[r] = qself;
    }
#pragma unroll
    for (int j = 0; j < kFewMax; ++j)
      if (j < a.few_n && live) qv[j] = (NORM && j == a.self) ? qself : static_cast<const T*>(a.few_row[j])[r];
    const long long s0 = a.slice_ptr[slice], s1 = a.slice_ptr[slice + 1];
    const int width = (int)((s1 - s0) / 32);
    const int32_t* colp = a.col + s0 + lane;
    const T* valp = static_cast<const T*>(a.val) + s0 + lane;
    T acc0 = T(0), acc1 = T(0);
    int c[W], cn[W];
    T v[W], vn[W];
    auto load_chunk = [&](int k0, int (&cc)[W], T (&vv)[W]) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const bool ok = k0 + j < width;
        const int off = (ok ? k0 + j : width - 1) * 32;  // clamped: a valid column, value 0
        cc[j] = ld_stream_col(colp + off);
        vv[j] = ok ? step::ld_stream(valp + off) : T(0);
      }
    };
    if (width > 0) load_chunk(0, cn, vn);
    for (int k0 = 0; k0 < width; k0 += W) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        c[j] = cn[j];
        v[j] = vn[j];
      }
      T g[W];
#pragma unroll
      for (int j = 0; j < W; ++j) g[j] = __ldg(x + c[j]);
      if (k0 + W < width) load_chunk(k0 + W, cn, vn);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const T gv = NORM ? g[j] * inv : g[j];
        if (j & 1)
          acc1 = fma(v[j], gv, acc1);
        else
          acc0 = fma(v[j], gv, acc0);
      }
    }
    if (live) {
      yv = acc0 + acc1;
      static_cast<T*>(a.y)[r] = yv;
    }
  }
  // this block's share of <few_j, y>: products in fp64, lanes and warps added in a fixed order
#pragma unroll
  for (int j = 0; j < kFewMax; ++j) {
    if (j < a.few_n) {
      const double w = warp_sum(static_cast<double>(qv[j]) * static_cast<double>(yv));
      if (lane == 0) share_s[j][warp] = w;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < a.few_n) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += share_s[threadIdx.x][w];
    a.partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

}  // namespace bl
