// Tensor-core (tcgen05 / TMEM) sweep of the matrix-free Gram operator, fp32 data.
//
// Reference behaviour replaced: the pairwise squared distance of
// /root/reference/src/matfree_extensions/util/gp_util.py:87-95 (Matern) and :168-176 (RBF),
//     s2_ij = max(0, |x_i|^2 + |x_j|^2 - 2 x_i.x_j),
// evaluated for a [128 x 256] tile of pairs by ONE group of `tcgen05.mma.kind::tf32` instructions.
//
// TF32 keeps 11 significand bits, the tolerance of this path is 1e-5, so every fp32 coordinate is
// split x = hi + lo (both exactly representable in TF32) and the contraction runs over the
// concatenated slots
//     A row i : [ hi(x) | lo(x) | hi(x) | 1     | 1      | 1      | 0.. ]
//     B row j : [ hi(y) | hi(y) | lo(y) | b_hi  | b_mid  | b_lo   | 0.. ]     b = -|y_j|^2 / 2 (exact 3-way split)
// so that the fp32 accumulator in TMEM holds  x.y - |y|^2/2  with an error of ~3 * 2^-24 |x||y| (the
// dropped lo*lo term and the last bit of lo) -- the level of an fp32 dot product.  The row term
// -|x_i|^2/2 is a per-thread constant of the epilogue (thread = row) and is folded into its first FMA,
// and the diagonal i == j is set to s2 = 0 exactly, as the fp32 evaluation of the reference's
// expression gives.  3d + 3 slots, padded to a multiple of 8 (one MMA consumes 8 TF32 slots):
// d = 9 -> 32 slots -> 4 MMAs of 128x256x8 per tile.
//
// Operands live in global memory in the canonical K-major no-swizzle core-matrix order
// ([16-byte slot chunk][point][4 floats]); a tile is a handful of contiguous 1-D TMA bulk copies and
// is consumed straight from shared memory by the tensor core.  Warp roles (one CTA per SM):
//   warp 0      TMA producer  (A tile once; B tile + v / q / x^T tile per stage)
//   warp 1      TMEM allocation + single-thread MMA issue, tcgen05.commit -> mbarriers
//   warps 2..9  epilogue: tcgen05.ld the accumulator (thread = row i, registers = columns j),
//               sqrt / exp on the MUFU, k_ij v_j accumulated per row.  Two accumulator buffers
//               (2 x 256 TMEM columns) let the MMA of tile t+1 run under the epilogue of tile t.
// The kernel is bound by the MUFU epilogue (rsqrt + ex2 per pair), not by the tensor pipe: the MMA
// group of a tile takes ~512 clocks, its epilogue ~4096.
#pragma once

#include "operators.cuh"
#include "tma_pipeline.cuh"

namespace bl {
namespace gramtc {

constexpr int kM = 128;          // rows per CTA  (UMMA M)
constexpr int kN = 256;          // columns per tile (UMMA N)
constexpr int kThreads = 320;    // 10 warps
constexpr int kEpiWarps = 8;
constexpr int kTmemCols = 512;   // two fp32 accumulators of kN columns
constexpr int kMaxSlots = 64;    // 3d + 3 <= 64  ->  d <= 20
constexpr int kXtRows = 20;      // rows of the transposed coordinate tile (adjoint sweep), >= d

__host__ __device__ inline int slots_for(int d) { return (3 * d + 3 + 7) / 8 * 8; }

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(slot_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// canonical layout ((8, m), 2) : ((16 B, SBO), LBO) -- LBO = distance between the two 16-byte slot
// chunks of one MMA, SBO = distance between groups of 8 rows.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(bar))
               : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane (blocking)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqf(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- operand packing ------------------------------------------------------------------------
// xs [n][dp] (scaled inputs), xx [n] (squared norms, same values as the ALU kernel uses)
//   -> opA / opB [slots/4][npad][4]   (points >= n are zero rows)
//   -> xt [kXtRows][npad]             (transposed fp32 coordinates for the adjoint epilogue)
__global__ void k_gram_tc_pack(int64_t n, int64_t npad, int d, int dp, int slots, const float* __restrict__ xs,
                               const float* __restrict__ xx, float* __restrict__ opA, float* __restrict__ opB,
                               float* __restrict__ xt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = i < n;
    float a[kMaxSlots], b[kMaxSlots];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) a[s] = b[s] = 0.f;
    if (live) {
      for (int k = 0; k < d; ++k) {
        const float x = xs[i * dp + k];
        const float hi = __uint_as_float(to_tf32(x));
        const float lo = __uint_as_float(to_tf32(x - hi));
        a[k] = hi, a[d + k] = lo, a[2 * d + k] = hi;
        b[k] = hi, b[d + k] = hi, b[2 * d + k] = lo;
      }
      const float h = -0.5f * xx[i];
      const float h0 = __uint_as_float(to_tf32(h));
      const float h1 = __uint_as_float(to_tf32(h - h0));
      const float h2 = __uint_as_float(to_tf32((h - h0) - h1));  // <= 2 bits left: exact
      a[3 * d] = 1.f, a[3 * d + 1] = 1.f, a[3 * d + 2] = 1.f;
      b[3 * d] = h0, b[3 * d + 1] = h1, b[3 * d + 2] = h2;
    }
    for (int c = 0; c < slots / 4; ++c) {
      float4 va, vb;
      va.x = a[4 * c], va.y = a[4 * c + 1], va.z = a[4 * c + 2], va.w = a[4 * c + 3];
      vb.x = b[4 * c], vb.y = b[4 * c + 1], vb.z = b[4 * c + 2], vb.w = b[4 * c + 3];
      reinterpret_cast<float4*>(opA)[(int64_t)c * npad + i] = va;
      reinterpret_cast<float4*>(opB)[(int64_t)c * npad + i] = vb;
    }
    for (int k = 0; k < kXtRows; ++k) xt[(int64_t)k * npad + i] = (live && k < d) ? xs[i * dp + k] : 0.f;
  }
}

// ---- shared-memory plan ---------------------------------------------------------------------
struct Plan {
  int ksteps;   // MMAs per tile (slots / 8)
  int stages;   // B-tile ring depth
  uint32_t a_bytes, b_bytes, aux_bytes, stage_bytes, bar_off, total;
};
// aux per stage: `aux_rows` rows of kN floats
__host__ __device__ inline Plan make_plan_rows(int slots, int aux_rows) {
  Plan p;
  p.ksteps = slots / 8;
  p.a_bytes = (uint32_t)slots * kM * 4;
  p.b_bytes = (uint32_t)slots * kN * 4;
  p.aux_bytes = (uint32_t)kN * 4 * aux_rows;
  p.stage_bytes = p.b_bytes + p.aux_bytes;
  const uint32_t budget = 220 * 1024;
  p.stages = 3;
  while (p.stages > 2 && p.a_bytes + p.stages * p.stage_bytes > budget) --p.stages;
  p.bar_off = p.a_bytes + p.stages * p.stage_bytes;
  p.total = p.bar_off + 256;
  if (p.total < 120 * 1024) p.total = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  return p;
}
// sweeps: v [kN] (+ q [kN] + x^T [kXtRows][kN] for the adjoint sweep)
__host__ __device__ inline Plan make_plan(int slots, bool adj) { return make_plan_rows(slots, adj ? 2 + kXtRows : 1); }
constexpr int kBatchMax = 16;  // (lambda, q) pairs per pass of the batched parameter-cotangent sweep
__host__ __device__ inline Plan make_plan_batch(int slots) { return make_plan_rows(slots, kBatchMax + kXtRows); }

// (unscaled) kernel values from the accumulator acc = x.y - |y|^2/2, two entries at a time.
//   KIND 0: Matern-3/2 (1+s) e^{-s}   KIND 1: Matern-1/2 e^{-s}   KIND 2: RBF e^{-s2/2}
// crow: per-row constant, (|x_i|^2 + eps) for the Matern kinds, (-|x_i|^2 / 2) for RBF.
struct Eval2 {
  float2 k;   // kernel value / sigma
  float2 e;   // e^{-s}  (Matern) or k (RBF)
  float2 ri;  // 1 / s   (Matern)
  bool pos0, pos1;  // s2 > 0 before the clamp (the clamp has zero derivative)   gp_util.py:92-95
};
template <int KIND>
__device__ __forceinline__ Eval2 kernel_from_acc(float2 acc, float2 crow) {
  Eval2 o;
  if (KIND == 2) {
    float2 a = __fadd2_rn(acc, crow);  // -s2/2
    o.pos0 = a.x < 0.f, o.pos1 = a.y < 0.f;
    a.x = fminf(a.x, 0.f), a.y = fminf(a.y, 0.f);  // clamp s2 at zero
    const float2 t = __fmul2_rn(a, make_float2(1.4426950408889634f, 1.4426950408889634f));
    o.k = make_float2(ex2f(t.x), ex2f(t.y));
    o.e = o.k;
    o.ri = make_float2(0.f, 0.f);
    return o;
  }
  const float eps = 1.1920928955078125e-07f;
  float2 t = __ffma2_rn(acc, make_float2(-2.f, -2.f), crow);  // s2 + eps
  o.pos0 = t.x > eps, o.pos1 = t.y > eps;
  t.x = fmaxf(t.x, eps), t.y = fmaxf(t.y, eps);  // max(s2, 0) + eps
  o.ri = make_float2(rsqf(t.x), rsqf(t.y));
  const float2 s = __fmul2_rn(t, o.ri);
  const float2 u = __fmul2_rn(s, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  o.e = make_float2(ex2f(u.x), ex2f(u.y));
  o.k = KIND == 0 ? __ffma2_rn(s, o.e, o.e) : o.e;
  return o;
}
// d k / d s2 (divided by sigma), zero where the clamp is active
template <int KIND>
__device__ __forceinline__ float2 dkernel_from_eval(const Eval2& o) {
  float2 dk = __fmul2_rn(KIND == 2 ? o.k : o.e, make_float2(-0.5f, -0.5f));
  if (KIND == 1) dk = __fmul2_rn(dk, o.ri);
  dk.x = o.pos0 ? dk.x : 0.f;
  dk.y = o.pos1 ? dk.y : 0.f;
  return dk;
}

// grid = (row tiles of 128, column splits).
//   ADJ == false: part[split][i] = sigma sum_{j in split} k_ij v_j
//   ADJ == true : the same with v = lam, plus per-CTA partial sums (layout of the ALU kernel)
//                 gpart[cta][d]  = sum lam_i q_j k_ij               (k includes sigma)
//                 gpart[cta][k]  = sum lam_i q_j dk_ij (x_ik - x_jk)^2,  k < d
// dbg != nullptr: CTA (dbg_bx, dbg_by) writes the raw accumulator (x.y - |y|^2/2) of its first tile to dbg[128][256].
template <int KIND, bool ADJ, int D>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_sweep(int64_t n, int64_t npad, int d, int slots, const float* __restrict__ opA,
                const float* __restrict__ opB, const float* __restrict__ xt, const float* __restrict__ xx,
                const float* __restrict__ consts,
                const float* __restrict__ v, const float* __restrict__ q, float* __restrict__ part,
                double* __restrict__ gpart, float* __restrict__ dbg, int dbg_bx, int dbg_by) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Plan pl = make_plan(slots, ADJ);
  uint8_t* smA = smem;
  uint8_t* smS = smem + pl.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
  uint64_t* full = bars;            // [stages]  TMA -> MMA, epilogue
  uint64_t* empty = bars + 3;       // [stages]  MMA commit + 8 epilogue warps -> TMA
  uint64_t* acc_full = bars + 6;    // [2]       MMA commit -> epilogue
  uint64_t* acc_empty = bars + 8;   // [2]       8 epilogue warps -> MMA
  uint64_t* a_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  __shared__ double ycomb[kM];
  __shared__ double gred[kEpiWarps][kXtRows + 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * kM;
  const int64_t tiles_total = (n + kN - 1) / kN;
  const int64_t per = (tiles_total + gridDim.y - 1) / gridDim.y;
  const int64_t t0 = per * blockIdx.y;
  const int64_t t1 = t0 + per < tiles_total ? t0 + per : tiles_total;
  const int ntiles = t0 < t1 ? (int)(t1 - t0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < pl.stages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, 1 + kEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      tma::mbar_init(acc_full + b, 1);
      tma::mbar_init(acc_empty + b, kEpiWarps);
    }
    tma::mbar_init(a_full, 1);
    tma::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && ntiles > 0) {
      tma::mbar_arrive_expect_tx(a_full, pl.a_bytes);
      for (int c = 0; c < 2 * pl.ksteps; ++c)
        tma::bulk_g2s(smA + (size_t)c * kM * 16, opA + ((int64_t)c * npad + i0) * 4, kM * 16, a_full);
    }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      tma::mbar_wait(empty + s, ph ^ 1u);
      uint8_t* smB = smS + (size_t)s * pl.stage_bytes;
      float* smV = reinterpret_cast<float*>(smB + pl.b_bytes);
      float* smQ = smV + kN;
      float* smX = smV + 2 * kN;
      const int64_t jt = (t0 + it) * kN;
      const int w = (int)((n - jt) < kN ? (n - jt) : kN);
      const int wv = w & ~3;  // whole 16-byte granules by TMA, the ragged end by this warp
      if (w < kN) {
        for (int c = wv + lane; c < kN; c += 32) {
          smV[c] = c < w ? v[jt + c] : 0.f;
          if (ADJ) smQ[c] = c < w ? q[jt + c] : 0.f;
        }
        __syncwarp();
      }
      if (lane == 0) {
        tma::mbar_arrive_expect_tx(full + s, pl.b_bytes + (uint32_t)wv * 4 * (ADJ ? 2 : 1) + (ADJ ? D * kN * 4 : 0));
        for (int c = 0; c < 2 * pl.ksteps; ++c)
          tma::bulk_g2s(smB + (size_t)c * kN * 16, opB + ((int64_t)c * npad + jt) * 4, kN * 16, full + s);
        if (wv > 0) tma::bulk_g2s(smV, v + jt, (uint32_t)wv * 4, full + s);
        if (ADJ) {
          if (wv > 0) tma::bulk_g2s(smQ, q + jt, (uint32_t)wv * 4, full + s);
          for (int k = 0; k < D; ++k) tma::bulk_g2s(smX + (size_t)k * kN, xt + (int64_t)k * npad + jt, kN * 4, full + s);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issue =====
    const uint32_t idesc = instr_desc(kM, kN);
    if (ntiles > 0) tma::mbar_wait(a_full, 0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_empty + b, bph ^ 1u);
      fence_after_sync();
      if (lane == 0) {
        const uint32_t sa = tma::smem_u32(smA), sb = tma::smem_u32(smS + (size_t)s * pl.stage_bytes);
        for (int k = 0; k < pl.ksteps; ++k) {
          const uint64_t da = smem_desc(sa + (uint32_t)k * 2 * kM * 16, kM * 16, 128);
          const uint64_t db = smem_desc(sb + (uint32_t)k * 2 * kN * 16, kN * 16, 128);
          mma_tf32(tmem_base + (uint32_t)b * kN, da, db, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(empty + s);
        mma_commit(acc_full + b);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread = row (TMEM lane), registers = columns =====
    const int qd = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // column half of the tile
    const int row = qd * 32 + lane;
    const bool live = i0 + row < n;
    const float xxi = live ? xx[i0 + row] : 0.f;
    const float cr = KIND == 2 ? -0.5f * xxi : xxi + 1.1920928955078125e-07f;
    const float2 crow = make_float2(cr, cr);
    const uint32_t acc_diag = __float_as_uint(0.5f * xxi);  // accumulator value that makes s2 exactly zero
    double yacc = 0.0, uacc = 0.0, dacc[ADJ ? D : 1];
    float2 nxi[ADJ ? D : 1];  // (-x_ik, -x_ik)
    if (ADJ) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const float x = xt[(int64_t)k * npad + i0 + row];
        nxi[k] = make_float2(-x, -x);
        dacc[k] = 0.0;
      }
    }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const float* smV = reinterpret_cast<const float*>(smS + (size_t)s * pl.stage_bytes + pl.b_bytes);
      const float* smQ = smV + kN;
      const float* smX = smV + 2 * kN;
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_full + b, bph);
      fence_after_sync();
      const int64_t jt = (t0 + it) * kN;
      const bool diag_tile = jt < i0 + kM && i0 < jt + kN;  // the tile holds pairs with i == j
      float2 y0 = make_float2(0.f, 0.f), y1 = y0, u0 = y0, u1 = y0, dl[ADJ ? D : 1];
      if (ADJ) {
#pragma unroll
        for (int k = 0; k < D; ++k) dl[k] = make_float2(0.f, 0.f);
      }
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = half * 128 + cc * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(b * kN + col0), r);
        if (dbg != nullptr && it == 0 && (int)blockIdx.x == dbg_bx && (int)blockIdx.y == dbg_by) {
#pragma unroll
          for (int c = 0; c < 32; ++c) dbg[(size_t)row * kN + col0 + c] = __uint_as_float(r[c]);
        }
        if (diag_tile) {
          const int jd = (int)(i0 + row - jt) - col0;  // column of this chunk with j == i (if any)
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = c == jd ? acc_diag : r[c];
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float4 v4 = *reinterpret_cast<const float4*>(smV + col0 + 4 * p);
          const float2 a01 = make_float2(__uint_as_float(r[4 * p]), __uint_as_float(r[4 * p + 1]));
          const float2 a23 = make_float2(__uint_as_float(r[4 * p + 2]), __uint_as_float(r[4 * p + 3]));
          const Eval2 e01 = kernel_from_acc<KIND>(a01, crow), e23 = kernel_from_acc<KIND>(a23, crow);
          y0 = __ffma2_rn(e01.k, make_float2(v4.x, v4.y), y0);
          y1 = __ffma2_rn(e23.k, make_float2(v4.z, v4.w), y1);
          if (ADJ) {
            const float4 q4 = *reinterpret_cast<const float4*>(smQ + col0 + 4 * p);
            const float2 q01 = make_float2(q4.x, q4.y), q23 = make_float2(q4.z, q4.w);
            u0 = __ffma2_rn(e01.k, q01, u0);
            u1 = __ffma2_rn(e23.k, q23, u1);
            const float2 g01 = __fmul2_rn(dkernel_from_eval<KIND>(e01), q01);
            const float2 g23 = __fmul2_rn(dkernel_from_eval<KIND>(e23), q23);
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const float4 x4 = *reinterpret_cast<const float4*>(smX + (size_t)k * kN + col0 + 4 * p);
              const float2 d01 = __fadd2_rn(make_float2(x4.x, x4.y), nxi[k]);
              const float2 d23 = __fadd2_rn(make_float2(x4.z, x4.w), nxi[k]);
              dl[k] = __ffma2_rn(g01, __fmul2_rn(d01, d01), dl[k]);
              dl[k] = __ffma2_rn(g23, __fmul2_rn(d23, d23), dl[k]);
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tma::mbar_arrive(acc_empty + b);
        tma::mbar_arrive(empty + s);
      }
      yacc += (double)(y0.x + y0.y) + (double)(y1.x + y1.y);
      if (ADJ) {
        uacc += (double)(u0.x + u0.y) + (double)(u1.x + u1.y);
#pragma unroll
        for (int k = 0; k < D; ++k) dacc[k] += (double)(dl[k].x + dl[k].y);
      }
    }
    // combine the two column halves, scale, store
    const double sigma = (double)consts[0];
    if (half == 1) ycomb[row] = yacc;
    tma::named_bar_sync(1, kEpiWarps * 32);
    if (half == 0 && live) part[(int64_t)blockIdx.y * n + i0 + row] = (float)((yacc + ycomb[row]) * sigma);
    if (ADJ) {
      const double li = live ? (double)v[i0 + row] * sigma : 0.0;
      const int ew = warp - 2;
      double t = warp_sum(li * uacc);
      if (lane == 0) gred[ew][D] = t;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        t = warp_sum(li * dacc[k]);
        if (lane == 0) gred[ew][k] = t;
      }
      tma::named_bar_sync(1, kEpiWarps * 32);
      if (warp == 2 && lane <= D) {
        double sacc = 0.0;
        for (int e = 0; e < kEpiWarps; ++e) sacc += gred[e][lane];
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        if (lane == D)
          gpart[blk * (d + 1) + d] = sacc;
        else if (lane < d)
          gpart[blk * (d + 1) + lane] = sacc;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// Deferred parameter cotangent of M <= kBatchMax matvec VJPs at once (the adjoint sweep of the Krylov
// loops defers them: arnoldi.py:207-209 only needs A^T lambda inside the loop):
//     sum_m d<lam_m, K(theta) q_m>/dtheta = sum_ij W_ij dk_ij/dtheta,   W_ij = sum_m lam_m[i] q_m[j],
// so the kernel tile (distances on the tensor pipe, sqrt / exp on the MUFU) and the per-dimension
// (x_ik - x_jk)^2 accumulation are paid ONCE for the M pairs; only the rank-M weight costs M FMAs per
// kernel entry.  Same warp roles and pipeline as k_gram_tc_sweep; per-CTA partial sums in gpart.
template <int KIND, int D>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_gradbatch(int64_t n, int64_t npad, int d, int slots, const float* __restrict__ opA,
                    const float* __restrict__ opB, const float* __restrict__ xt, const float* __restrict__ xx,
                    const float* __restrict__ consts, const float* __restrict__ Qrows, int64_t ldq,
                    const float* __restrict__ Lrows, int64_t ldl, int M, double* __restrict__ gpart) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Plan pl = make_plan_batch(slots);
  uint8_t* smA = smem;
  uint8_t* smS = smem + pl.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
  uint64_t* full = bars;
  uint64_t* empty = bars + 3;
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 8;
  uint64_t* a_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  __shared__ double gred[kEpiWarps][kXtRows + 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * kM;
  const int64_t tiles_total = (n + kN - 1) / kN;
  const int64_t per = (tiles_total + gridDim.y - 1) / gridDim.y;
  const int64_t t0 = per * blockIdx.y;
  const int64_t t1 = t0 + per < tiles_total ? t0 + per : tiles_total;
  const int ntiles = t0 < t1 ? (int)(t1 - t0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < pl.stages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, 1 + kEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      tma::mbar_init(acc_full + b, 1);
      tma::mbar_init(acc_empty + b, kEpiWarps);
    }
    tma::mbar_init(a_full, 1);
    tma::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && ntiles > 0) {
      tma::mbar_arrive_expect_tx(a_full, pl.a_bytes);
      for (int c = 0; c < 2 * pl.ksteps; ++c)
        tma::bulk_g2s(smA + (size_t)c * kM * 16, opA + ((int64_t)c * npad + i0) * 4, kM * 16, a_full);
    }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      tma::mbar_wait(empty + s, ph ^ 1u);
      uint8_t* smB = smS + (size_t)s * pl.stage_bytes;
      float* smQ = reinterpret_cast<float*>(smB + pl.b_bytes);  // [kBatchMax][kN]
      float* smX = smQ + kBatchMax * kN;                        // [kXtRows][kN]
      const int64_t jt = (t0 + it) * kN;
      // basis rows are zero-padded up to their stride: copy what the row holds, zero the rest of the tile
      const int w = (int)((ldq - jt) < kN ? (ldq - jt) : kN);
      const int wv = w & ~3;
      if (wv < kN) {
        for (int m = 0; m < M; ++m)
          for (int c = wv + lane; c < kN; c += 32) smQ[m * kN + c] = (jt + c < n) ? Qrows[(int64_t)m * ldq + jt + c] : 0.f;
        __syncwarp();
      }
      if (lane == 0) {
        tma::mbar_arrive_expect_tx(full + s, pl.b_bytes + (uint32_t)M * wv * 4 + (uint32_t)D * kN * 4);
        for (int c = 0; c < 2 * pl.ksteps; ++c)
          tma::bulk_g2s(smB + (size_t)c * kN * 16, opB + ((int64_t)c * npad + jt) * 4, kN * 16, full + s);
        if (wv > 0)
          for (int m = 0; m < M; ++m) tma::bulk_g2s(smQ + (size_t)m * kN, Qrows + (int64_t)m * ldq + jt, (uint32_t)wv * 4, full + s);
        for (int k = 0; k < D; ++k) tma::bulk_g2s(smX + (size_t)k * kN, xt + (int64_t)k * npad + jt, kN * 4, full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issue =====
    const uint32_t idesc = instr_desc(kM, kN);
    if (ntiles > 0) tma::mbar_wait(a_full, 0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_empty + b, bph ^ 1u);
      fence_after_sync();
      if (lane == 0) {
        const uint32_t sa = tma::smem_u32(smA), sb = tma::smem_u32(smS + (size_t)s * pl.stage_bytes);
        for (int k = 0; k < pl.ksteps; ++k) {
          const uint64_t da = smem_desc(sa + (uint32_t)k * 2 * kM * 16, kM * 16, 128);
          const uint64_t db = smem_desc(sb + (uint32_t)k * 2 * kN * 16, kN * 16, 128);
          mma_tf32(tmem_base + (uint32_t)b * kN, da, db, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(empty + s);
        mma_commit(acc_full + b);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const bool live = i0 + row < n;
    const float xxi = live ? xx[i0 + row] : 0.f;
    const float cr = KIND == 2 ? -0.5f * xxi : xxi + 1.1920928955078125e-07f;
    const float2 crow = make_float2(cr, cr);
    const uint32_t acc_diag = __float_as_uint(0.5f * xxi);
    double uacc = 0.0, dacc[D];
    float2 nxi[D], lam2[kBatchMax];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float x = xt[(int64_t)k * npad + i0 + row];
      nxi[k] = make_float2(-x, -x);
      dacc[k] = 0.0;
    }
#pragma unroll
    for (int m = 0; m < kBatchMax; ++m) {
      const float l = (live && m < M) ? Lrows[(int64_t)m * ldl + i0 + row] : 0.f;
      lam2[m] = make_float2(l, l);
    }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const float* smQ = reinterpret_cast<const float*>(smS + (size_t)s * pl.stage_bytes + pl.b_bytes);
      const float* smX = smQ + kBatchMax * kN;
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_full + b, bph);
      fence_after_sync();
      const int64_t jt = (t0 + it) * kN;
      const bool diag_tile = jt < i0 + kM && i0 < jt + kN;
      float2 u0 = make_float2(0.f, 0.f), u1 = u0, dl[D];
#pragma unroll
      for (int k = 0; k < D; ++k) dl[k] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = half * 128 + cc * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(b * kN + col0), r);
        if (diag_tile) {
          const int jd = (int)(i0 + row - jt) - col0;
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = c == jd ? acc_diag : r[c];
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float2 a01 = make_float2(__uint_as_float(r[4 * p]), __uint_as_float(r[4 * p + 1]));
          const float2 a23 = make_float2(__uint_as_float(r[4 * p + 2]), __uint_as_float(r[4 * p + 3]));
          const Eval2 e01 = kernel_from_acc<KIND>(a01, crow), e23 = kernel_from_acc<KIND>(a23, crow);
          float2 w01 = make_float2(0.f, 0.f), w23 = w01;  // W_ij = sum_m lam_m[i] q_m[j]
#pragma unroll
          for (int m = 0; m < kBatchMax; ++m) {
            if (m < M) {
              const float4 q4 = *reinterpret_cast<const float4*>(smQ + (size_t)m * kN + col0 + 4 * p);
              w01 = __ffma2_rn(lam2[m], make_float2(q4.x, q4.y), w01);
              w23 = __ffma2_rn(lam2[m], make_float2(q4.z, q4.w), w23);
            }
          }
          u0 = __ffma2_rn(e01.k, w01, u0);
          u1 = __ffma2_rn(e23.k, w23, u1);
          const float2 g01 = __fmul2_rn(dkernel_from_eval<KIND>(e01), w01);
          const float2 g23 = __fmul2_rn(dkernel_from_eval<KIND>(e23), w23);
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const float4 x4 = *reinterpret_cast<const float4*>(smX + (size_t)k * kN + col0 + 4 * p);
            const float2 d01 = __fadd2_rn(make_float2(x4.x, x4.y), nxi[k]);
            const float2 d23 = __fadd2_rn(make_float2(x4.z, x4.w), nxi[k]);
            dl[k] = __ffma2_rn(g01, __fmul2_rn(d01, d01), dl[k]);
            dl[k] = __ffma2_rn(g23, __fmul2_rn(d23, d23), dl[k]);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tma::mbar_arrive(acc_empty + b);
        tma::mbar_arrive(empty + s);
      }
      uacc += (double)(u0.x + u0.y) + (double)(u1.x + u1.y);
#pragma unroll
      for (int k = 0; k < D; ++k) dacc[k] += (double)(dl[k].x + dl[k].y);
    }
    const double sigma = (double)consts[0];
    const int ew = warp - 2;
    double t = warp_sum(sigma * uacc);
    if (lane == 0) gred[ew][D] = t;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      t = warp_sum(sigma * dacc[k]);
      if (lane == 0) gred[ew][k] = t;
    }
    tma::named_bar_sync(1, kEpiWarps * 32);
    if (warp == 2 && lane <= D) {
      double sacc = 0.0;
      for (int e = 0; e < kEpiWarps; ++e) sacc += gred[e][lane];
      const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      if (lane == D)
        gpart[blk * (d + 1) + d] = sacc;
      else if (lane < d)
        gpart[blk * (d + 1) + lane] = sacc;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// Matvec of P <= kBatchMax vectors at once (lockstep Krylov runs over probes): every kernel tile --
// distances on the tensor pipe, sqrt / exp on the MUFU -- is evaluated ONCE and applied to the P
// vectors (P extra FMAs per entry), so P matvecs cost about as much as one.
//   part[split][p][i] = sigma sum_{j in split} k_ij v_p[j]
struct VecPtrs {
  const float* p[kBatchMax];
};
__host__ __device__ inline Plan make_plan_multi(int slots) { return make_plan_rows(slots, kBatchMax); }

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
k_gram_tc_multi(int64_t n, int64_t npad, int slots, const float* __restrict__ opA, const float* __restrict__ opB,
                const float* __restrict__ xx, const float* __restrict__ consts, VecPtrs vecs, int P,
                float* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Plan pl = make_plan_multi(slots);
  uint8_t* smA = smem;
  uint8_t* smS = smem + pl.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bar_off);
  uint64_t* full = bars;
  uint64_t* empty = bars + 3;
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 8;
  uint64_t* a_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  __shared__ double ycomb[kBatchMax][kM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * kM;
  const int64_t tiles_total = (n + kN - 1) / kN;
  const int64_t per = (tiles_total + gridDim.y - 1) / gridDim.y;
  const int64_t t0 = per * blockIdx.y;
  const int64_t t1 = t0 + per < tiles_total ? t0 + per : tiles_total;
  const int ntiles = t0 < t1 ? (int)(t1 - t0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < pl.stages; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, 1 + kEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      tma::mbar_init(acc_full + b, 1);
      tma::mbar_init(acc_empty + b, kEpiWarps);
    }
    tma::mbar_init(a_full, 1);
    tma::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && ntiles > 0) {
      tma::mbar_arrive_expect_tx(a_full, pl.a_bytes);
      for (int c = 0; c < 2 * pl.ksteps; ++c)
        tma::bulk_g2s(smA + (size_t)c * kM * 16, opA + ((int64_t)c * npad + i0) * 4, kM * 16, a_full);
    }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      tma::mbar_wait(empty + s, ph ^ 1u);
      uint8_t* smB = smS + (size_t)s * pl.stage_bytes;
      float* smV = reinterpret_cast<float*>(smB + pl.b_bytes);  // [kBatchMax][kN]
      const int64_t jt = (t0 + it) * kN;
      const int w = (int)((n - jt) < kN ? (n - jt) : kN);
      const int wv = w & ~3;
      if (w < kN) {
        for (int p = 0; p < P; ++p)
          for (int c = wv + lane; c < kN; c += 32) smV[p * kN + c] = c < w ? vecs.p[p][jt + c] : 0.f;
        __syncwarp();
      }
      if (lane == 0) {
        tma::mbar_arrive_expect_tx(full + s, pl.b_bytes + (uint32_t)P * wv * 4);
        for (int c = 0; c < 2 * pl.ksteps; ++c)
          tma::bulk_g2s(smB + (size_t)c * kN * 16, opB + ((int64_t)c * npad + jt) * 4, kN * 16, full + s);
        if (wv > 0)
          for (int p = 0; p < P; ++p) tma::bulk_g2s(smV + (size_t)p * kN, vecs.p[p] + jt, (uint32_t)wv * 4, full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issue =====
    const uint32_t idesc = instr_desc(kM, kN);
    if (ntiles > 0) tma::mbar_wait(a_full, 0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_empty + b, bph ^ 1u);
      fence_after_sync();
      if (lane == 0) {
        const uint32_t sa = tma::smem_u32(smA), sb = tma::smem_u32(smS + (size_t)s * pl.stage_bytes);
        for (int k = 0; k < pl.ksteps; ++k) {
          const uint64_t da = smem_desc(sa + (uint32_t)k * 2 * kM * 16, kM * 16, 128);
          const uint64_t db = smem_desc(sb + (uint32_t)k * 2 * kN * 16, kN * 16, 128);
          mma_tf32(tmem_base + (uint32_t)b * kN, da, db, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(empty + s);
        mma_commit(acc_full + b);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const bool live = i0 + row < n;
    const float xxi = live ? xx[i0 + row] : 0.f;
    const float cr = KIND == 2 ? -0.5f * xxi : xxi + 1.1920928955078125e-07f;
    const float2 crow = make_float2(cr, cr);
    const uint32_t acc_diag = __float_as_uint(0.5f * xxi);
    double yacc[kBatchMax];
#pragma unroll
    for (int p = 0; p < kBatchMax; ++p) yacc[p] = 0.0;
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % pl.stages;
      const uint32_t ph = (uint32_t)(it / pl.stages) & 1u;
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const float* smV = reinterpret_cast<const float*>(smS + (size_t)s * pl.stage_bytes + pl.b_bytes);
      tma::mbar_wait(full + s, ph);
      tma::mbar_wait(acc_full + b, bph);
      fence_after_sync();
      const int64_t jt = (t0 + it) * kN;
      const bool diag_tile = jt < i0 + kM && i0 < jt + kN;
      float2 y[kBatchMax];
#pragma unroll
      for (int p = 0; p < kBatchMax; ++p) y[p] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = half * 128 + cc * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(b * kN + col0), r);
        if (diag_tile) {
          const int jd = (int)(i0 + row - jt) - col0;
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = c == jd ? acc_diag : r[c];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 a01 = make_float2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]));
          const float2 a23 = make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
          const Eval2 e01 = kernel_from_acc<KIND>(a01, crow), e23 = kernel_from_acc<KIND>(a23, crow);
#pragma unroll
          for (int p = 0; p < kBatchMax; ++p) {
            if (p < P) {
              const float4 v4 = *reinterpret_cast<const float4*>(smV + (size_t)p * kN + col0 + 4 * q);
              y[p] = __ffma2_rn(e01.k, make_float2(v4.x, v4.y), y[p]);
              y[p] = __ffma2_rn(e23.k, make_float2(v4.z, v4.w), y[p]);
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tma::mbar_arrive(acc_empty + b);
        tma::mbar_arrive(empty + s);
      }
#pragma unroll
      for (int p = 0; p < kBatchMax; ++p) yacc[p] += (double)(y[p].x + y[p].y);
    }
    const double sigma = (double)consts[0];
    if (half == 1) {
#pragma unroll
      for (int p = 0; p < kBatchMax; ++p) ycomb[p][row] = yacc[p];
    }
    tma::named_bar_sync(1, kEpiWarps * 32);
    if (half == 0 && live) {
#pragma unroll
      for (int p = 0; p < kBatchMax; ++p)
        if (p < P) part[((int64_t)blockIdx.y * P + p) * n + i0 + row] = (float)((yacc[p] + ycomb[p][row]) * sigma);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace gramtc
}  // namespace bl
