// Dense operand of the reference's unit tests and the user-callback operand.
//   mode 0: matvec = p @ s            (/root/reference/tests/test_arnoldi/test_hessenberg_forward.py:20)
//   mode 1: matvec = (p + p.T) @ s    (/root/reference/tests/test_lanczos/test_tridiag_adjoint.py:20-21)
// Small n only (BASELINE config 1: n = 100); one warp per output row.
#include "operators.cuh"

namespace bl {
namespace {

template <typename T>
__global__ void k_dense_matvec(int64_t n, int mode, bool transpose, const T* __restrict__ P,
                               const T* __restrict__ x, T* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  double acc = 0.0;
  for (int64_t c = lane; c < n; c += 32) {
    T a;
    if (mode == 1)
      a = P[r * n + c] + P[c * n + r];
    else
      a = transpose ? P[c * n + r] : P[r * n + c];
    acc += static_cast<double>(a * x[c]);
  }
  acc = warp_sum(acc);
  if (lane == 0) y[r] = static_cast<T>(acc);
}

// grad[r][c] += lam[r] q[c]  (+ q[r] lam[c] for the symmetrised operand)
template <typename T>
__global__ void k_dense_outer(int64_t n, int mode, const T* __restrict__ q, const T* __restrict__ lam,
                              T* __restrict__ grad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const int64_t r = e / n, c = e % n;
  T g = lam[r] * q[c];
  if (mode == 1) g += q[r] * lam[c];
  grad[e] += g;
}

// out[j] = M[j, idx]  /  out[j] = M[j, j]   (M = P, or P + P^T for the symmetrised operand)
template <typename T>
__global__ void k_dense_elements(int64_t n, int mode, const T* __restrict__ P, const int64_t* __restrict__ index,
                                 T* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t c = index ? index[0] : j;
  out[j] = mode == 1 ? P[j * n + c] + P[c * n + j] : P[j * n + c];
}

}  // namespace

struct DenseOperator : bl_operator {
  int mode = 0;
  const void* P = nullptr;
  int bound_dtype = -1;
  DevBuf grad;

  int num_params() const override { return 1; }
  int64_t param_size(int) const override { return n * n; }
  int set_params(int dtype, const void* const* params, int num, cudaStream_t) override {
    BL_REQUIRE(num == 1 && params && params[0], "dense operator takes one parameter (n x n)");
    P = params[0];
    bound_dtype = dtype;
    return grad.ensure((size_t)n * n * dtype_size(dtype));
  }
  template <typename T>
  int mv(bool transpose, const void* x, void* y, cudaStream_t s) {
    const int blocks = (int)((n * 32 + 255) / 256);
    k_dense_matvec<T><<<blocks, 256, 0, s>>>(n, mode, transpose, static_cast<const T*>(P), static_cast<const T*>(x), static_cast<T*>(y));
    BL_LAUNCHED();
    return BL_OK;
  }
  int matvec(int dtype, const void* x, void* y, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    return dtype == BL_F32 ? mv<float>(false, x, y, s) : mv<double>(false, x, y, s);
  }
  int elements(int dtype, const int64_t* index, void* out, cudaStream_t s) {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    const int blocks = (int)((n + 255) / 256);
    if (dtype == BL_F32)
      k_dense_elements<float><<<blocks, 256, 0, s>>>(n, mode, static_cast<const float*>(P), index, static_cast<float*>(out));
    else
      k_dense_elements<double><<<blocks, 256, 0, s>>>(n, mode, static_cast<const double*>(P), index, static_cast<double*>(out));
    BL_LAUNCHED();
    return BL_OK;
  }
  int element_diagonal(int dtype, void* out, cudaStream_t s) override { return elements(dtype, nullptr, out, s); }
  int element_column(int dtype, const int64_t* index, void* out, cudaStream_t s) override {
    return elements(dtype, index, out, s);
  }
  int vjp(int dtype, const void* q, const void* lam, void* z, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    if (z) BL_CHECK(dtype == BL_F32 ? mv<float>(true, lam, z, s) : mv<double>(true, lam, z, s));
    const int blocks = (int)((n * n + 255) / 256);
    if (dtype == BL_F32)
      k_dense_outer<float><<<blocks, 256, 0, s>>>(n, mode, static_cast<const float*>(q), static_cast<const float*>(lam), grad.as<float>());
    else
      k_dense_outer<double><<<blocks, 256, 0, s>>>(n, mode, static_cast<const double*>(q), static_cast<const double*>(lam), grad.as<double>());
    BL_LAUNCHED();
    return BL_OK;
  }
  int grad_zero(int dtype, cudaStream_t s) override {
    BL_CHECK(grad.ensure((size_t)n * n * dtype_size(dtype)));
    BL_CUDA(cudaMemsetAsync(grad.p, 0, (size_t)n * n * dtype_size(dtype), s));
    return BL_OK;
  }
  int grad_export(int dtype, void* const* grads, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 1 && grads && grads[0], "dense operator has one gradient buffer");
    BL_CUDA(cudaMemcpyAsync(grads[0], grad.p, (size_t)n * n * dtype_size(dtype), cudaMemcpyDeviceToDevice, s));
    return BL_OK;
  }
};

// The reference's arbitrary user callable: the host layer enqueues the work itself.
struct CallbackOperator : bl_operator {
  bl_matvec_cb mv = nullptr;
  bl_vjp_cb vj = nullptr;
  void* user = nullptr;
  int num_params() const override { return 0; }
  int64_t param_size(int) const override { return 0; }
  int set_params(int, const void* const*, int, cudaStream_t) override { return BL_OK; }
  int matvec(int dtype, const void* x, void* y, cudaStream_t s) override {
    if (mv(user, dtype, x, y, s) != 0) {
      set_error("user matvec callback failed");
      return BL_ECALLBACK;
    }
    return BL_OK;
  }
  int vjp(int dtype, const void* q, const void* lam, void* z, cudaStream_t s) override {
    BL_REQUIRE(vj != nullptr, "callback operator has no vjp");
    if (vj(user, dtype, q, lam, z, s) != 0) {
      set_error("user vjp callback failed");
      return BL_ECALLBACK;
    }
    return BL_OK;
  }
  int grad_zero(int, cudaStream_t) override { return BL_OK; }
  int grad_export(int, void* const*, int, cudaStream_t) override { return BL_OK; }
};

}  // namespace bl

extern "C" {

int bl_op_dense_create(int64_t n, int mode, bl_operator_t** op) {
  BL_REQUIRE(op != nullptr && n > 0 && (mode == 0 || mode == 1), "bad dense operator arguments");
  auto* o = new bl::DenseOperator();
  o->n = n;
  o->mode = mode;
  *op = o;
  return BL_OK;
}

int bl_op_callback_create(int64_t n, bl_matvec_cb matvec_cb, bl_vjp_cb vjp_cb, void* user,
                          bl_operator_t** op) {
  BL_REQUIRE(op != nullptr && n > 0 && matvec_cb != nullptr, "bad callback operator arguments");
  auto* o = new bl::CallbackOperator();
  o->n = n;
  o->mv = matvec_cb;
  o->vj = vjp_cb;
  o->user = user;
  *op = o;
  return BL_OK;
}

int bl_op_destroy(bl_operator_t* op) {
  delete op;
  return BL_OK;
}
int bl_op_size(const bl_operator_t* op, int64_t* n) {
  BL_REQUIRE(op && n, "NULL argument");
  *n = op->n;
  return BL_OK;
}
int bl_op_num_params(const bl_operator_t* op, int* num) {
  BL_REQUIRE(op && num, "NULL argument");
  *num = op->num_params();
  return BL_OK;
}
int bl_op_param_size(const bl_operator_t* op, int index, int64_t* numel) {
  BL_REQUIRE(op && numel && index >= 0 && index < op->num_params(), "bad parameter index");
  *numel = op->param_size(index);
  return BL_OK;
}
int bl_op_set_params(bl_operator_t* op, int dtype, const void* const* params, int num, void* stream) {
  BL_REQUIRE(op && (dtype == BL_F32 || dtype == BL_F64), "bad operator/dtype");
  return op->set_params(dtype, params, num, bl::as_stream(stream));
}
int bl_op_matvec(bl_operator_t* op, int dtype, const void* x, void* y, void* stream) {
  BL_REQUIRE(op && x && y && (dtype == BL_F32 || dtype == BL_F64), "bad matvec arguments");
  return op->matvec(dtype, x, y, bl::as_stream(stream));
}
int bl_op_vjp(bl_operator_t* op, int dtype, const void* q, const void* lam, void* z, void* stream) {
  BL_REQUIRE(op && q && lam && (dtype == BL_F32 || dtype == BL_F64), "bad vjp arguments");
  return op->vjp(dtype, q, lam, z, bl::as_stream(stream));
}
int bl_op_grad_zero(bl_operator_t* op, int dtype, void* stream) {
  BL_REQUIRE(op && (dtype == BL_F32 || dtype == BL_F64), "bad operator/dtype");
  return op->grad_zero(dtype, bl::as_stream(stream));
}
int bl_op_grad_export(bl_operator_t* op, int dtype, void* const* grads, int num, void* stream) {
  BL_REQUIRE(op && (dtype == BL_F32 || dtype == BL_F64), "bad operator/dtype");
  return op->grad_export(dtype, grads, num, bl::as_stream(stream));
}

}  // extern "C"
