// Operator interface behind bl_operator_t: the reference's user-supplied `matvec(v, *params)`
// and the pullback `jax.vjp` gives the adjoint sweep (arnoldi.py:207-209).
#pragma once

#include "common.cuh"

namespace bl {
// SELL-32 operand in device memory (layout: sparse.cu).  What the operator phase of k_step_tma reads when the
// operator call of a Krylov step rides in the step kernel.
struct SellView {
  const int64_t* slice_ptr = nullptr;  // nslices + 1
  const int32_t* col = nullptr;
  const void* val = nullptr;           // T[nslots], T = the bound dtype
  int64_t nslices = 0, nrows = 0;
};
}  // namespace bl

struct bl_operator {
  int64_t n = 0;  // square operators: length of x and y
  virtual ~bl_operator() {}
  virtual int num_params() const = 0;
  virtual int64_t param_size(int index) const = 0;
  virtual int set_params(int dtype, const void* const* params, int num, cudaStream_t s) = 0;
  virtual int matvec(int dtype, const void* x, void* y, cudaStream_t s) = 0;
  // z = A^T lam (skipped when z == nullptr); grad += d<lam, A(q)>/dparams
  virtual int vjp(int dtype, const void* q, const void* lam, void* z, cudaStream_t s) = 0;
  virtual int grad_zero(int dtype, cudaStream_t s) = 0;
  virtual int grad_export(int dtype, void* const* grads, int num, cudaStream_t s) = 0;
  // ALGORITHMIC bytes of one matvec / one vjp (roofline report); default: vectors only
  virtual double matvec_bytes(int dtype) const { return 2.0 * n * (dtype == BL_F32 ? 4 : 8); }
  // ... of one batched call over `count` vectors (operators that read their own data once per batch override)
  virtual double matvec_batch_bytes(int dtype, int count) const { return matvec_bytes(dtype) * count; }
  virtual double vjp_bytes(int dtype) const { return 3.0 * n * (dtype == BL_F32 ? 4 : 8); }
  // Forward step of the Krylov loops: q = v / *len (true division, arnoldi.py:80-81; the row is written up
  // to n_pad entries with zero padding) followed by y = A q.  Operators that can normalise on the fly
  // return BL_OK and save one pass over the vector; the default returns BL_EUNSUPPORTED (-1) and the
  // caller runs the two steps separately.
  virtual int matvec_normalised(int /*dtype*/, const void* /*v*/, const double* /*len*/, void* /*q_out*/,
                                int64_t /*n_pad*/, void* /*y*/, cudaStream_t) {
    return -1;
  }
  // Deferred parameter cotangent (the adjoint sweeps only need A^T lam inside the loop): operators that
  // return true provide `apply_transpose` (z = A^T lam, no gradient work) and `vjp_batch`
  // (grad += sum_m d<lam_m, A(q_m)>/dparams for `count` rows of two row-strided arrays) -- for the Gram
  // operator one batched sweep costs about as much as ONE per-step cotangent sweep.
  virtual bool deferred_grad(int /*dtype*/) const { return false; }
  // ALGORITHMIC bytes of the two halves of a deferred cotangent (roofline report)
  virtual double apply_transpose_bytes(int dtype) const { return vjp_bytes(dtype); }
  virtual double vjp_batch_bytes(int dtype, int /*count*/) const { return vjp_bytes(dtype); }
  virtual int apply_transpose(int /*dtype*/, const void* /*lam*/, void* /*z*/, cudaStream_t) {
    bl::set_error("apply_transpose is not implemented for this operator");
    return BL_EINVAL;
  }
  virtual int vjp_batch(int dtype, const void* Q, int64_t ldq, const void* Lam, int64_t ldl, int count,
                        cudaStream_t s) {
    const size_t w = dtype == BL_F32 ? 4 : 8;
    for (int m = 0; m < count; ++m) {
      const int rc = vjp(dtype, static_cast<const char*>(Q) + (size_t)m * ldq * w,
                         static_cast<const char*>(Lam) + (size_t)m * ldl * w, nullptr, s);
      if (rc != BL_OK) return rc;
    }
    return BL_OK;
  }
  // Several independent vectors at once (lockstep Krylov runs over probes): out[p] = A in[p].  The
  // default loops; the Gram operator evaluates each kernel tile once for all of them.
  virtual int matvec_batch(int dtype, int count, const void* const* in, void* const* out, cudaStream_t s) {
    for (int p = 0; p < count; ++p) {
      const int rc = matvec(dtype, in[p], out[p], s);
      if (rc != BL_OK) return rc;
    }
    return BL_OK;
  }
  // matvec_normalised for `count` lockstep runs: q_out[p] = v[p] / *len[p], y[p] = A q_out[p].  -1: not supported
  // (the caller normalises separately and calls matvec_batch).
  virtual int matvec_normalised_batch(int /*dtype*/, int /*count*/, const void* const* /*v*/, const double* const* /*len*/,
                                      void* const* /*q_out*/, int64_t /*n_pad*/, void* const* /*y*/, cudaStream_t) {
    return -1;
  }
  virtual int apply_transpose_batch(int dtype, int count, const void* const* in, void* const* out, cudaStream_t s) {
    for (int p = 0; p < count; ++p) {
      const int rc = apply_transpose(dtype, in[p], out[p], s);
      if (rc != BL_OK) return rc;
    }
    return BL_OK;
  }
  // Operators stored as SELL-32 expose A (transpose = false) or A^T (true) to the step kernel; others return false.
  virtual bool sell_view(int /*dtype*/, bool /*transpose*/, bl::SellView* /*out*/) const { return false; }
  // Lazily evaluated matrix elements (the `lazy_kernel(i, j)` of gp_util.py:257-258 / the
  // `matrix_element` callback of low_rank.py): diagonal and one column, for the partial Cholesky.
  // For the Gram operator these are the KERNEL entries (no noise term), as in the reference.
  virtual int element_diagonal(int /*dtype*/, void* /*out*/, cudaStream_t) {
    bl::set_error("this operator does not expose matrix elements");
    return BL_EINVAL;
  }
  virtual int element_column(int /*dtype*/, const int64_t* /*index_dev*/, void* /*out*/, cudaStream_t) {
    bl::set_error("this operator does not expose matrix elements");
    return BL_EINVAL;
  }
};

namespace bl {

// Small RAII device buffer used by operators for their own (long-lived) storage.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int ensure(size_t want) {
    if (want <= bytes) return BL_OK;
    release();
    if (cudaMalloc(&p, want ? want : 1) != cudaSuccess) {
      set_error("cudaMalloc failed in operator storage");
      p = nullptr;
      return BL_ENOMEM;
    }
    bytes = want;
    return BL_OK;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

}  // namespace bl
