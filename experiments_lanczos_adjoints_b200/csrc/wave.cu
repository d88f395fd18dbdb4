// Wave-equation stencil operand and its adjoint.
//
// Reference behaviour replaced: `pde_wave_anisotropic(...).rhs` with `boundary_neumann`
// (/root/reference/src/matfree_extensions/util/pde_util.py:126-157) and its `jax.vjp`:
//   state x = (u, du), each g x g, raveled to 2 g^2;   A x = (du, scale^2 * conv(stencil, pad_edge(u)))
// `convolve2d(stencil, padded, "valid")` is a true convolution:
//   conv(u)[i,j] = sum_{a,b} st[a,b] * u[clamp(i+1-a), clamp(j+1-b)]     (edge replicate = index clamp)
// HBM-bound: ~ (3 reads + 2 writes) * g^2 values per matvec; one thread per grid point, rows of
// the grid are contiguous so every access is coalesced.
#include "operators.cuh"

namespace bl {
namespace {

struct Stencil {
  double w[9];
};

__device__ __forceinline__ int64_t clampi(int64_t v, int64_t hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

template <typename T>
__device__ __forceinline__ T conv_at(const T* __restrict__ u, int64_t g, int64_t i, int64_t j, const Stencil& st) {
  T s = T(0);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const T w = static_cast<T>(st.w[a * 3 + b]);
      if (w != T(0)) s = fma(w, u[clampi(i + 1 - a, g - 1) * g + clampi(j + 1 - b, g - 1)], s);
    }
  return s;
}

template <typename T>
__global__ void k_wave_matvec(int64_t g, Stencil st, const T* __restrict__ scale, const T* __restrict__ x,
                              T* __restrict__ y) {
  // 2-D launch: blockIdx.y = grid row, x = column (no 64-bit div/mod per point; rows are contiguous)
  const int64_t gg = g * g;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = blockIdx.y; i < g; i += gridDim.y) {
    if (j >= g) continue;
    const int64_t p = i * g + j;
    y[p] = x[gg + p];  // d/dt u = du
    const T sc = scale[p];
    y[gg + p] = conv_at<T>(x, g, i, j, st) * (sc * sc);  // fx * constrain(scale), constrain = square
  }
}

// transpose of conv (with the clamp folded in): contributions to cell (p, q) come from
//   i = p + a - 1 (if inside), plus i = 0 when p == 0 and a == 2, plus i = g-1 when p == g-1 and a == 0
template <typename T>
__device__ __forceinline__ T conv_t_at(const T* __restrict__ w, int64_t g, int64_t p, int64_t q, const Stencil& st) {
  T s = T(0);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int64_t is[2];
    int ni = 0;
    const int64_t i0 = p + a - 1;
    if (i0 >= 0 && i0 <= g - 1) is[ni++] = i0;
    if (p == 0 && a == 2) is[ni++] = 0;
    if (p == g - 1 && a == 0) is[ni++] = g - 1;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const T wt = static_cast<T>(st.w[a * 3 + b]);
      if (wt == T(0)) continue;
      int64_t js[2];
      int nj = 0;
      const int64_t jj0 = q + b - 1;
      if (jj0 >= 0 && jj0 <= g - 1) js[nj++] = jj0;
      if (q == 0 && b == 2) js[nj++] = 0;
      if (q == g - 1 && b == 0) js[nj++] = g - 1;
      for (int ii = 0; ii < ni; ++ii)
        for (int jj = 0; jj < nj; ++jj) s = fma(wt, w[is[ii] * g + js[jj]], s);
    }
  }
  return s;
}

// tmp = scale^2 * lam_du ; grad += 2 scale lam_du conv(q_u) ; z_du = lam_u
template <typename T>
__global__ void k_wave_vjp_a(int64_t g, Stencil st, const T* __restrict__ scale, const T* __restrict__ q,
                             const T* __restrict__ lam, T* __restrict__ tmp, T* __restrict__ z, T* __restrict__ grad) {
  const int64_t gg = g * g;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = blockIdx.y; i < g; i += gridDim.y) {
    if (j >= g) continue;
    const int64_t p = i * g + j;
    const T sc = scale[p], ld = lam[gg + p];
    tmp[p] = sc * sc * ld;
    grad[p] = fma(T(2) * sc * ld, conv_at<T>(q, g, i, j, st), grad[p]);
    if (z) z[gg + p] = lam[p];
  }
}

// z_u = conv^T(tmp)
template <typename T>
__global__ void k_wave_vjp_b(int64_t g, Stencil st, const T* __restrict__ tmp, T* __restrict__ z) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = blockIdx.y; i < g; i += gridDim.y)
    if (j < g) z[i * g + j] = conv_t_at<T>(tmp, g, i, j, st);
}

}  // namespace

struct WaveOperator : bl_operator {
  int64_t g = 0;
  Stencil st;
  const void* scale = nullptr;
  int bound_dtype = -1;
  DevBuf grad, tmp;

  int num_params() const override { return 1; }
  int64_t param_size(int) const override { return g * g; }
  dim3 blocks() const { return dim3((unsigned)((g + 255) / 256), (unsigned)std::min<int64_t>(g, 65535)); }

  int set_params(int dtype, const void* const* params, int num, cudaStream_t) override {
    BL_REQUIRE(num == 1 && params && params[0], "wave operator takes one parameter (scale, g x g)");
    scale = params[0];
    bound_dtype = dtype;
    BL_CHECK(grad.ensure((size_t)g * g * dtype_size(dtype)));
    return tmp.ensure((size_t)g * g * dtype_size(dtype));
  }
  int matvec(int dtype, const void* x, void* y, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    if (dtype == BL_F32)
      k_wave_matvec<float><<<blocks(), 256, 0, s>>>(g, st, (const float*)scale, (const float*)x, (float*)y);
    else
      k_wave_matvec<double><<<blocks(), 256, 0, s>>>(g, st, (const double*)scale, (const double*)x, (double*)y);
    BL_LAUNCHED();
    return BL_OK;
  }
  template <typename T>
  int vjp_t(const T* q, const T* lam, T* z, cudaStream_t s) {
    k_wave_vjp_a<T><<<blocks(), 256, 0, s>>>(g, st, (const T*)scale, q, lam, tmp.as<T>(), z, grad.as<T>());
    BL_LAUNCHED();
    if (z) {
      k_wave_vjp_b<T><<<blocks(), 256, 0, s>>>(g, st, tmp.as<T>(), z);
      BL_LAUNCHED();
    }
    return BL_OK;
  }
  int vjp(int dtype, const void* q, const void* lam, void* z, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    return dtype == BL_F32 ? vjp_t<float>((const float*)q, (const float*)lam, (float*)z, s)
                           : vjp_t<double>((const double*)q, (const double*)lam, (double*)z, s);
  }
  int grad_zero(int dtype, cudaStream_t s) override {
    BL_CHECK(grad.ensure((size_t)g * g * dtype_size(dtype)));
    BL_CUDA(cudaMemsetAsync(grad.p, 0, (size_t)g * g * dtype_size(dtype), s));
    return BL_OK;
  }
  int grad_export(int dtype, void* const* grads, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 1 && grads && grads[0], "wave operator has one gradient buffer");
    BL_CUDA(cudaMemcpyAsync(grads[0], grad.p, (size_t)g * g * dtype_size(dtype), cudaMemcpyDeviceToDevice, s));
    return BL_OK;
  }
};

}  // namespace bl

extern "C" int bl_op_wave_create(int64_t grid, const double* stencil3x3_host, bl_operator_t** op) {
  BL_REQUIRE(op && stencil3x3_host && grid >= 2, "bad wave operator arguments");
  auto* o = new bl::WaveOperator();
  o->g = grid;
  o->n = 2 * grid * grid;
  for (int k = 0; k < 9; ++k) o->st.w[k] = stencil3x3_host[k];
  *op = o;
  return BL_OK;
}
