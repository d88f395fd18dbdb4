// Streaming kernels of the Krylov loops (sm_100a).
//
// Everything the Arnoldi / Lanczos forward and adjoint sweeps do with the basis is one of
//   dots    : red[j] = <row_j, x>              for a block of basis rows      ("Q^T v")
//   combine : out = s * (sum_k a_k vec_k + sum_j c_j row_j)   (+ ||out||^2)   ("v - Q c")
// Both end with a grid-wide reduction whose result feeds the next kernel; the last block to
// finish reduces the per-block partials in a fixed order (deterministic, no float atomics)
// and runs a small "epilogue" that turns the reduced numbers into the scalars / coefficient
// vectors the next kernel reads from device memory — the host never synchronises.
#pragma once

#include "common.cuh"
#include "dist.cuh"

namespace bl {

constexpr int kDotsThreads = 256;     // 8 warps: warp w owns rows w, w+8, ...
constexpr int kCombineThreads = 128;
constexpr int kMaxVecTerms = 5;

// ---- epilogues --------------------------------------------------------------------------
enum EpiMode : int {
  EPI_NONE = 0,
  EPI_STORE,        // out_t[j] = red[j]                          (plain helper calls)
  EPI_INIT_NORM,    // len0 = sqrt(red[0]); inv_len = 1/len0; c = 1/len0
  EPI_FWD_A,        // H[j,i] = red[j]; coef[j] = H[j,i]                      arnoldi.py:87,99
  EPI_FWD_B,        // coef[j] = red[j] (+ H[j,i] = red[j] for the rows j < j0 the first pass skipped)  arnoldi.py:92
  EPI_FWD_NORM,     // len = sqrt(red[0]); H[i+1,i] = len; inv_len = 1/len    arnoldi.py:95-98
  EPI_ADJ_ETA,      // eta[j] = dH[j,K-1] - red[j]                            arnoldi.py:119
  EPI_ADJ_REPROJ,   // coef[j] = dH[j,idx] - red[j]   (j <= idx+1)            arnoldi.py:202-204
  EPI_ADJ_GAMMA,    // Gamma row, (Gamma+Gamma^T) row, beta_plus, alpha, 1/beta_minus  :212-219
  EPI_L3_ALPHA,     // alpha_i = red[0]; coefficients of the 3-term residual  lanczos.py:280-282
  EPI_L3_BETA,      // beta_i = sqrt(red[0]); inv_len = 1/beta_i              lanczos.py:283-284
  EPI_L3_ADJ_MUNU,  // mu, nu and the coefficients of lambda                  lanczos.py:322-325
  EPI_L3_ADJ_DOT,   // scal[slot] = red[0]
  EPI_L3_ADJ_FINAL  // coefficients of grad_initvec                           lanczos.py:311
};

// Slots of the device-resident double "scalar block".
enum ScalSlot : int {
  S_LEN = 0, S_INV_LEN, S_C, S_ALPHA, S_BETA_MINUS, S_ETA_IDX, S_NEG_ALPHA, S_TMP0, S_TMP1,
  S_MU, S_NU, S_B, S_INV_B, S_NEG_B_NU, S_A, S_DOT0, S_COUNT
};

struct Epi {
  int mode = EPI_NONE;
  int i = 0;      // forward step / adjoint idx
  int K = 0;      // krylov depth
  int m = 0;      // number of reduced values
  double* red = nullptr;       // reduced values (m doubles)
  double* scal = nullptr;      // scalar block
  double* coef = nullptr;      // coefficient vector written for the next combine
  double* coef2 = nullptr;     // second coefficient vector (beta_plus / Lambda rows)
  void* H = nullptr;           // K x K (dtype T), written by the forward
  const void* Hc = nullptr;    // K x K (dtype T), read by the adjoint
  const void* dH = nullptr;    // K x K (dtype T)
  double* Gamma = nullptr;     // K x K doubles
  const double* PiGamma = nullptr;  // K x K doubles
  double* eta = nullptr;       // K doubles
  void* out_t = nullptr;       // dtype-T output (EPI_STORE, c, alphas, betas)
  void* out_t2 = nullptr;
  const void* in_t = nullptr;  // dtype-T inputs for the 3-term adjoint (dalphas)
  const void* in_t2 = nullptr; // (dbetas)
  const void* in_t3 = nullptr; // (alphas)
  const void* in_t4 = nullptr; // (betas)
  int slot = 0;
  int j0 = 0;  // EPI_ADJ_GAMMA / EPI_FWD_A: red[] holds the dots with rows j0.. only; Gamma[idx, j < j0] = 0, h[j < j0] = 0
  // row sharding over peer memory: the `peer_count` values of red[] are summed over the ranks (dist.cuh) by
  // the block that runs the epilogue, before the epilogue consumes them.  Only the own mailbox and the
  // sequence number travel with the kernel; the peer table sits in the mailbox header.
  unsigned char* peer_mail = nullptr;
  unsigned long long peer_seq = 0;
  int peer_count = 0;
};

// Runs after `red[0..m)` has been written and the `nt` participating threads (index t) have synchronised;
// `sync()` is their barrier (the whole block for the last-block pattern, the consumer threads of k_step_tma).
struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

template <typename T, typename Sync>
__device__ void run_epilogue_impl(const Epi& e, const double* red, const int t, const int nt, Sync sync) {
  const int K = e.K, i = e.i;
  switch (e.mode) {
    case EPI_NONE:
      break;
    case EPI_STORE: {
      T* o = static_cast<T*>(e.out_t);
      for (int j = t; j < e.m; j += nt) o[j] = static_cast<T>(red[j]);
    } break;
    case EPI_INIT_NORM: {
      if (t == 0) {
        // dtype-T arithmetic as the reference: sqrt(dot(v,v)), 1/len
        T len = sqrt(static_cast<T>(red[0]));
        e.scal[S_LEN] = static_cast<double>(len);
        e.scal[S_INV_LEN] = static_cast<double>(T(1) / len);
        if (e.out_t) static_cast<T*>(e.out_t)[0] = T(1) / len;
      }
    } break;
    case EPI_FWD_A: {
      T* H = static_cast<T*>(e.H);  // red[] holds the dots with rows j0..i; h[j < j0] = 0 (H starts as zeros)
      for (int j = t; j < e.j0 + e.m; j += nt) {
        T h = j >= e.j0 ? static_cast<T>(red[j - e.j0]) : T(0);
        if (j >= e.j0) H[(size_t)j * K + i] = h;
        e.coef[j] = static_cast<double>(h);
      }
    } break;
    case EPI_FWD_B: {
      for (int j = t; j < e.m; j += nt) e.coef[j] = static_cast<double>(static_cast<T>(red[j]));
      // BL_FWD_SYMMETRIC: the first pass skipped rows j < j0, so the second pass's coefficient q_j . v' IS
      // q_j . (A q_i) up to rounding (v' differs from A q_i along q_{i-1}, q_i only): it completes column i of H.
      // O(eps |A|) for a symmetric operand; the host reads these entries to detect a non-symmetric one.
      if (e.H != nullptr)
        for (int j = t; j < e.j0; j += nt) static_cast<T*>(e.H)[(size_t)j * K + i] = static_cast<T>(red[j]);
    } break;
    case EPI_FWD_NORM: {
      if (t == 0) {
        T len = sqrt(static_cast<T>(red[0]));
        e.scal[S_LEN] = static_cast<double>(len);
        e.scal[S_INV_LEN] = static_cast<double>(T(1) / len);
        if (i + 1 < K) static_cast<T*>(e.H)[(size_t)(i + 1) * K + i] = len;  // dropped at i+1 == K
      }
    } break;
    // The adjoint's epilogues read what EARLIER kernels of the sweep wrote (Gamma rows, eta) and run redundantly in
    // every block of k_step_tma: through L2 (`ld.global.cg` / `st.global.cg`), like everything else that crosses blocks.
    case EPI_ADJ_ETA: {
      const T* dH = static_cast<const T*>(e.dH);
      for (int j = t; j < K; j += nt) {
        double r = e.m > 0 ? red[j] : 0.0;
        e.eta[j] = static_cast<double>(static_cast<T>(static_cast<double>(__ldcg(dH + (size_t)j * K + (K - 1))) - r));
        e.coef[j] = e.eta[j];
      }
    } break;
    case EPI_ADJ_REPROJ: {
      const T* dH = static_cast<const T*>(e.dH);
      // rows j <= idx+1 of P are active; p = mask * dH[:, idx]
      for (int j = t; j < e.m; j += nt)
        e.coef[j] = static_cast<double>(static_cast<T>(static_cast<double>(__ldcg(dH + (size_t)j * K + i)) - red[j]));
    } break;
    case EPI_ADJ_GAMMA: {
      const T* H = static_cast<const T*>(e.Hc);
      const int idx = i;
      // Gamma[idx, j] = lower_mask[idx, j] * (Pi_gamma[idx, j] - (A^T lam)^T q_j),  j <= idx
      for (int j = t; j < K; j += nt) {
        double g = 0.0;
        if (j <= idx && j >= e.j0) {
          g = __ldcg(e.PiGamma + (size_t)idx * K + j) - red[j - e.j0];
          if (j == idx) g *= 0.5;
          g = static_cast<double>(static_cast<T>(g));
        }
        __stcg(e.Gamma + (size_t)idx * K + j, g);
      }
      __threadfence();
      sync();
      for (int j = t; j < K; j += nt) {
        // (Gamma + Gamma^T)[idx, j]
        e.coef[j] = static_cast<double>(
            static_cast<T>(__ldcg(e.Gamma + (size_t)idx * K + j) + __ldcg(e.Gamma + (size_t)j * K + idx)));
        // -beta_plus[j] = -H[idx, j] for j > idx (diagonal and sub-diagonal removed)
        e.coef2[j] = j > idx ? -static_cast<double>(__ldcg(H + (size_t)idx * K + j)) : 0.0;
      }
      if (t == 0) {
        e.scal[S_NEG_ALPHA] = -static_cast<double>(__ldcg(H + (size_t)idx * K + idx));
        double bm = idx == 0 ? 1.0 : static_cast<double>(__ldcg(H + (size_t)idx * K + idx - 1));
        e.scal[S_BETA_MINUS] = bm;  // combine divides by it (lambda_k /= beta_minus)
        e.scal[S_ETA_IDX] = __ldcg(e.eta + idx);
      }
    } break;
    case EPI_L3_ALPHA: {
      if (t == 0) {
        T a = static_cast<T>(red[0]);
        static_cast<T*>(e.out_t)[i] = a;
        // residual = A x_i - a x_i - b_{i-1} x_{i-1}: coefficients for rows (i-1, i) or (i)
        if (i == 0) {
          e.coef[0] = -static_cast<double>(a);
        } else {
          e.coef[0] = -e.scal[S_B];
          e.coef[1] = -static_cast<double>(a);
        }
      }
    } break;
    case EPI_L3_BETA: {
      if (t == 0) {
        T b = sqrt(static_cast<T>(red[0]));
        static_cast<T*>(e.out_t)[i] = b;
        e.scal[S_B] = static_cast<double>(b);
        e.scal[S_LEN] = static_cast<double>(b);
        e.scal[S_INV_LEN] = static_cast<double>(T(1) / b);
      }
    } break;
    case EPI_L3_ADJ_DOT: {
      if (t == 0) e.scal[e.slot] = red[0];
    } break;
    case EPI_L3_ADJ_MUNU: {
      // red[0] = x_k . xi, red[1] = x_{k+1} . xi  (xi not yet divided by b_k);
      // scal[S_DOT0] = lambda_plus . x_k
      if (t == 0) {
        const int k = i;
        double b = static_cast<double>(static_cast<const T*>(e.in_t4)[k]);
        double da = static_cast<double>(static_cast<const T*>(e.in_t)[k]);
        double db = static_cast<double>(static_cast<const T*>(e.in_t2)[k]);
        double a = static_cast<double>(static_cast<const T*>(e.in_t3)[k]);
        double inv_b = 1.0 / b;
        double mu = db - e.scal[S_DOT0] + red[1] * inv_b;
        double nu = da + red[0] * inv_b;
        mu = static_cast<double>(static_cast<T>(mu));
        nu = static_cast<double>(static_cast<T>(nu));
        e.scal[S_MU] = mu;
        e.scal[S_NU] = nu;
        e.scal[S_B] = b;
        e.scal[S_A] = a;
        e.scal[S_INV_B] = -inv_b;          // lambda = -xi/b + mu x_{k+1} + nu x_k
        e.scal[S_NEG_B_NU] = -b * nu;      // xi' = ... - b nu x_{k+1}
        e.coef[0] = nu;                    // rows (k, k+1) of xs
        e.coef[1] = mu;
        e.coef2[0] = 0.0;
        e.coef2[1] = -b * nu;
      }
    } break;
    case EPI_L3_ADJ_FINAL: {
      // grad_initvec = ((xi . x_0) x_0 - xi) / ||v||
      if (t == 0) {
        double vn = static_cast<double>(static_cast<const T*>(e.in_t)[0]);
        e.coef[0] = red[0] / vn;
        e.scal[S_TMP0] = -1.0 / vn;
      }
    } break;
  }
}

// Runs in the last block, after `red[0..m)` has been written by that same block and a __syncthreads().
template <typename T>
__device__ void run_epilogue(const Epi& e) {
  if (e.peer_mail != nullptr && e.peer_count > 0)
    dist::peer_allreduce_block(dist::view_from_mailbox(e.peer_mail, e.peer_seq), e.red, e.peer_count);
  run_epilogue_impl<T>(e, e.red, (int)threadIdx.x, (int)blockDim.x, BlockSync());
}

// ---- dots ---------------------------------------------------------------------------------
struct RowBlock {
  const void* base = nullptr;  // first row of the basis buffer
  long long ld = 0;            // row stride in elements
  int row0 = 0;                // first row of the block
  int nrows = 0;
  const double* coef = nullptr;  // combine: coefficient of row (row0 + j) is sign * coef[coef0 + j]
  int coef0 = 0;
  double sign = 1.0;
};

template <typename T>
__device__ __forceinline__ double dot_vec(const typename Vec<T>::type& a, const typename Vec<T>::type& b) {
  T x[Vec<T>::N], y[Vec<T>::N];
  vec_unpack(a, x);
  vec_unpack(b, y);
  T s = T(0);
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) s = fma(x[k], y[k], s);
  return static_cast<double>(s);
}

// partials layout: [row j][block b] so the final reduction reads contiguous memory per row.
template <typename T>
__global__ void __launch_bounds__(kDotsThreads)
k_dots(RowBlock blk, const T* __restrict__ x, long long n, double* __restrict__ partials,
       unsigned int* counter, Epi epi) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long ngroups = n / VN;  // full 16-byte groups
  // balanced contiguous chunk of groups per block
  const long long per = (ngroups + gridDim.x - 1) / gridDim.x;
  const long long g0 = per * blockIdx.x;
  long long g1 = g0 + per;
  if (g1 > ngroups) g1 = ngroups;
  const V* xv = reinterpret_cast<const V*>(x);
  const T* base = static_cast<const T*>(blk.base);

  for (int j = warp; j < blk.nrows; j += nwarps) {
    const V* row = reinterpret_cast<const V*>(base + (long long)(blk.row0 + j) * blk.ld);
    T acc[4] = {T(0), T(0), T(0), T(0)};
    long long g = g0 + lane;
    // 4 independent 128-bit loads in flight per lane
    for (; g + 96 < g1; g += 128) {
      V q0 = ld_stream(row + g), q1 = ld_stream(row + g + 32), q2 = ld_stream(row + g + 64),
        q3 = ld_stream(row + g + 96);
      V x0 = __ldg(xv + g), x1 = __ldg(xv + g + 32), x2 = __ldg(xv + g + 64), x3 = __ldg(xv + g + 96);
      T a[VN], b[VN];
      vec_unpack(q0, a); vec_unpack(x0, b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[0] = fma(a[k], b[k], acc[0]);
      vec_unpack(q1, a); vec_unpack(x1, b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[1] = fma(a[k], b[k], acc[1]);
      vec_unpack(q2, a); vec_unpack(x2, b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[2] = fma(a[k], b[k], acc[2]);
      vec_unpack(q3, a); vec_unpack(x3, b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[3] = fma(a[k], b[k], acc[3]);
    }
    for (; g < g1; g += 32) {
      V q0 = ld_stream(row + g);
      V x0 = __ldg(xv + g);
      T a[VN], b[VN];
      vec_unpack(q0, a); vec_unpack(x0, b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[0] = fma(a[k], b[k], acc[0]);
    }
    double s = (static_cast<double>(acc[0]) + static_cast<double>(acc[1])) +
               (static_cast<double>(acc[2]) + static_cast<double>(acc[3]));
    // scalar tail (n not a multiple of the vector width): last block, lane 0
    if (blockIdx.x == gridDim.x - 1 && lane == 0) {
      const T* rowt = base + (long long)(blk.row0 + j) * blk.ld;
      for (long long c = ngroups * VN; c < n; ++c) s += static_cast<double>(rowt[c] * x[c]);
    }
    s = warp_sum(s);
    if (lane == 0) partials[(size_t)j * gridDim.x + blockIdx.x] = s;
  }

  if (!last_block_done(counter)) return;
  // fixed-order reduction over blocks: one warp per row, lanes stride the blocks
  const int G = gridDim.x;
  for (int j = warp; j < blk.nrows; j += nwarps) {
    const double* p = partials + (size_t)j * G;
    double s = 0.0;
    for (int b = lane; b < G; b += 32) s += __ldcg(p + b);
    s = warp_sum(s);
    if (lane == 0) epi.red[j] = s;
  }
  __syncthreads();
  run_epilogue<T>(epi);
}

// ---- combine -------------------------------------------------------------------------------
struct VecTerm {
  const void* ptr = nullptr;
  const double* coef_ptr = nullptr;  // device scalar (may be null)
  double coef_imm = 1.0;             // coefficient = coef_imm * (coef_ptr ? *coef_ptr : 1)
};

struct CombineArgs {
  long long n = 0;
  void* out = nullptr;
  int nvec = 0;
  VecTerm vec[kMaxVecTerms];
  RowBlock blk[2];
  const double* out_div_ptr = nullptr;  // out /= *out_div_ptr
  const double* out_mul_ptr = nullptr;  // out *= *out_mul_ptr
  void* out2 = nullptr;                 // optional second copy of the result
  double* partials = nullptr;           // ||out||^2 partials (NORM)
  unsigned int* counter = nullptr;
  Epi epi;
};

template <typename T, bool NORM>
__global__ void __launch_bounds__(kCombineThreads)
k_combine(CombineArgs a) {
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  extern __shared__ unsigned char smem_raw[];
  T* coef = reinterpret_cast<T*>(smem_raw);  // nrows0 + nrows1 coefficients
  __shared__ double red_smem[32];
  __shared__ T vcoef[kMaxVecTerms];
  __shared__ T oscale[2];

  const int n0 = a.blk[0].nrows, n1 = a.blk[1].nrows;
  for (int j = threadIdx.x; j < n0 + n1; j += blockDim.x) {
    const RowBlock& b = j < n0 ? a.blk[0] : a.blk[1];
    const int jj = j < n0 ? j : j - n0;
    coef[j] = static_cast<T>(b.sign * b.coef[b.coef0 + jj]);
  }
  if (threadIdx.x < a.nvec) {
    const VecTerm& v = a.vec[threadIdx.x];
    vcoef[threadIdx.x] = static_cast<T>(v.coef_imm * (v.coef_ptr ? *v.coef_ptr : 1.0));
  }
  if (threadIdx.x == 0) {
    oscale[0] = a.out_mul_ptr ? static_cast<T>(*a.out_mul_ptr) : T(1);
    oscale[1] = a.out_div_ptr ? static_cast<T>(*a.out_div_ptr) : T(1);
  }
  __syncthreads();

  const long long ngroups = (a.n + VN - 1) / VN;
  const long long nfull = a.n / VN;
  double ss = 0.0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
       g += (long long)gridDim.x * blockDim.x) {
    T acc[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[k] = T(0);
    if (g < nfull) {
      for (int t = 0; t < a.nvec; ++t) {
        V v = reinterpret_cast<const V*>(a.vec[t].ptr)[g];  // may alias `out`: plain load
        T e[VN];
        vec_unpack(v, e);
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] = fma(vcoef[t], e[k], acc[k]);
      }
#pragma unroll
      for (int bi = 0; bi < 2; ++bi) {
        const RowBlock& b = a.blk[bi];
        const T* cf = coef + (bi == 0 ? 0 : n0);
        const V* row = reinterpret_cast<const V*>(static_cast<const T*>(b.base) + (long long)b.row0 * b.ld) + g;
        const long long ldv = b.ld / VN;
        int j = 0;
        for (; j + 8 <= b.nrows; j += 8) {
          V q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) q[u] = ld_stream(row + (long long)(j + u) * ldv);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            T e[VN];
            vec_unpack(q[u], e);
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[k] = fma(cf[j + u], e[k], acc[k]);
          }
        }
        for (; j < b.nrows; ++j) {
          V q = ld_stream(row + (long long)j * ldv);
          T e[VN];
          vec_unpack(q, e);
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[k] = fma(cf[j], e[k], acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        acc[k] = acc[k] * oscale[0] / oscale[1];
        if (NORM) ss += static_cast<double>(acc[k] * acc[k]);
      }
      reinterpret_cast<V*>(a.out)[g] = vec_pack(acc);
      if (a.out2) reinterpret_cast<V*>(a.out2)[g] = vec_pack(acc);
    } else {
      // scalar tail group
      for (long long c = g * VN; c < a.n; ++c) {
        T s = T(0);
        for (int t = 0; t < a.nvec; ++t) s = fma(vcoef[t], static_cast<const T*>(a.vec[t].ptr)[c], s);
        for (int bi = 0; bi < 2; ++bi) {
          const RowBlock& b = a.blk[bi];
          const T* cf = coef + (bi == 0 ? 0 : n0);
          const T* basep = static_cast<const T*>(b.base) + (long long)b.row0 * b.ld;
          for (int j = 0; j < b.nrows; ++j) s = fma(cf[j], basep[(long long)j * b.ld + c], s);
        }
        s = s * oscale[0] / oscale[1];
        if (NORM) ss += static_cast<double>(s * s);
        static_cast<T*>(a.out)[c] = s;
        if (a.out2) static_cast<T*>(a.out2)[c] = s;
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

// ---- small dense helpers -------------------------------------------------------------------
// Pi_gamma = -dc*c*e1 e1^T + H dH^T - G   (K x K doubles; G = dQ^T Q or absent)   arnoldi.py:127
template <typename T>
__global__ void k_pi_gamma(int K, const T* __restrict__ H, const T* __restrict__ dH, const T* dc,
                           const T* c, const double* __restrict__ G, double* __restrict__ out) {
  const int a = blockIdx.x, b = threadIdx.x + blockIdx.y * blockDim.x;
  if (b >= K) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k)
    s += static_cast<double>(H[(size_t)a * K + k]) * static_cast<double>(dH[(size_t)b * K + k]);
  if (a == 0 && b == 0 && dc) s -= static_cast<double>(dc[0]) * static_cast<double>(c[0]);
  if (G) s -= G[(size_t)a * K + b];
  out[(size_t)a * K + b] = s;
}

// G[a][b] = sum_col dQ[a][col] * Q[b][col]: each block owns a column slab and writes its own partial K x K (summed in a
// fixed order by k_gram_reduce).  One (Ka x Kb) block of the result per launch: rows a0.. of dQ against rows b0.. of Q
// (the host walks the blocks, at most 128 x 128 each, so that a thread owns at most two 4x4 patches whatever the depth).
// A tile is TK = 32 columns, staged TRANSPOSED in shared memory ([column][row], row stride kGramKP = 132): a warp loads 32
// consecutive columns of a row (one coalesced 128-byte request; the loads of a tile are issued before the first store),
// and a thread reads the four rows of its patch as ONE 16-byte word per operand and column -- 2 shared loads for 16 FMAs.
// The 32 products of a tile are summed in T, the tile sums in double (a conversion and an fp64 add per product made the
// F2D pipe the bound: 5.4 ms at K = 100, n = 1M, where the bytes take 0.15 ms).
constexpr int kGramKP = 132;
constexpr int kGramThreads = 512;

template <typename T>
__device__ __forceinline__ void ld4_shared(const T* p, T (&o)[4]);
template <>
__device__ __forceinline__ void ld4_shared<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
}
template <>
__device__ __forceinline__ void ld4_shared<double>(const double* p, double (&o)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  o[0] = a.x, o[1] = a.y, o[2] = b.x, o[3] = b.y;
}

template <typename T, int TK>
__global__ void __launch_bounds__(kGramThreads)
k_gram_partial(int Ka, int Kb, int ldp, long long n, const T* __restrict__ dQ, const T* __restrict__ Q, long long ld,
               double* __restrict__ partial /* [gridDim.x][ldp*ldp], offset to (a0, b0) */) {
  static_assert(TK == 32, "a warp loads one row of a tile");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);  // [TK][kGramKP]
  T* sB = sA + (size_t)TK * kGramKP;       // [TK][kGramKP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kGramThreads / 32;
  const int K = Ka > Kb ? Ka : Kb;
  const int K4 = (K + 3) / 4 * 4;          // rows K .. K4-1 are staged as zeros (the patches read whole 4-row words)
  const int P = (Kb + 3) / 4;              // patches per row of patches
  const int PA = (Ka + 3) / 4;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long c0 = per * blockIdx.x;
  long long c1 = c0 + per;
  if (c1 > n) c1 = n;
  const int npatch = PA * P;
  constexpr int MAXP = 2;  // 32 x 32 patches of a 128 x 128 block over 512 threads
  constexpr int RMAX = 128 / NW;  // rows a warp stages per tile
  double acc[MAXP][16];
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[p][k] = 0.0;
  for (long long c = c0; c < c1; c += TK) {
    const int w = (int)((c1 - c) < TK ? (c1 - c) : TK);
    T ra[RMAX], rb[RMAX];
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      const int r = warp + i * NW;
      ra[i] = (lane < w && r < Ka) ? dQ[(long long)r * ld + c + lane] : T(0);
      rb[i] = (lane < w && r < Kb) ? Q[(long long)r * ld + c + lane] : T(0);
    }
    __syncthreads();  // the previous tile has been consumed
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      const int r = warp + i * NW;
      if (r < K4) {
        sA[lane * kGramKP + r] = ra[i];
        sB[lane * kGramKP + r] = rb[i];
      }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
      const int patch = threadIdx.x + p * kGramThreads;
      if (patch >= npatch) break;
      const int pa = (patch / P) * 4, pb = (patch % P) * 4;
      T tile[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) tile[k] = T(0);
#pragma unroll 8
      for (int k = 0; k < TK; ++k) {
        T av[4], bv[4];
        ld4_shared<T>(sA + k * kGramKP + pa, av);
        ld4_shared<T>(sB + k * kGramKP + pb, bv);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) tile[u * 4 + v] = fma(av[u], bv[v], tile[u * 4 + v]);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[p][k] += static_cast<double>(tile[k]);
    }
  }
  double* out = partial + (size_t)blockIdx.x * ldp * ldp;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    const int patch = threadIdx.x + p * kGramThreads;
    if (patch >= npatch) break;
    const int pa = (patch / P) * 4, pb = (patch % P) * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (pa + u < Ka && pb + v < Kb) out[(size_t)(pa + u) * ldp + pb + v] = acc[p][u * 4 + v];
  }
}

__global__ void k_gram_reduce(int KK, int nparts, const double* __restrict__ partial, double* __restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= KK) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * KK + e];
  G[e] = s;
}

// out[0..n) = x * mul / div ; out[n..n_pad) = 0  (keeps the row padding of a basis buffer
// clean).  `v /= length` is a true division as in the reference (arnoldi.py:80).
template <typename T>
__global__ void k_scale_copy(long long n, const T* __restrict__ x, double mul_imm, const double* div_ptr,
                             T* __restrict__ out, long long n_pad) {
  const T m = static_cast<T>(mul_imm);
  const T d = div_ptr ? static_cast<T>(*div_ptr) : T(1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 8;  // grid strides per round: their loads are issued before the first store
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_pad; i0 += U * stride) {
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      v[u] = i < n ? x[i] : T(0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n_pad) out[i] = i < n ? (div_ptr ? v[u] * m / d : v[u] * m) : T(0);
    }
  }
}

template <typename T>
__global__ void k_load_scalar(const T* src, double* dst) { *dst = static_cast<double>(*src); }

// Runs an epilogue on its own (no streaming pass in front of it).
template <typename T>
__global__ void k_epilogue_only(Epi epi) { run_epilogue<T>(epi); }

}  // namespace bl
