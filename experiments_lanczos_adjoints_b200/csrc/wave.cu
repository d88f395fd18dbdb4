// Wave-equation stencil operand and its adjoint.
//
// Reference behaviour replaced: `pde_wave_anisotropic(...).rhs` with `boundary_neumann`
// (/root/reference/src/matfree_extensions/util/pde_util.py:126-157) and its `jax.vjp`:
//   state x = (u, du), each g x g, raveled to 2 g^2;   A x = (du, scale^2 * conv(stencil, pad_edge(u)))
// `convolve2d(stencil, padded, "valid")` is a true convolution:
//   conv(u)[i,j] = sum_{a,b} st[a,b] * u[clamp(i+1-a), clamp(j+1-b)]     (edge replicate = index clamp)
// HBM-bound: ~ (3 reads + 2 writes) * g^2 values per matvec; one thread per grid point, rows of
// the grid are contiguous so every access is coalesced.
#include "dist.cuh"
#include "operators.cuh"

// Row-sharded use (one slab of grid rows per GPU): the operator covers `gy` rows x `gx` columns and
// may have a neighbour above / below.  Rows next to a neighbour read one halo row that the host
// layer exchanges (NCCL send/recv) into the operator's halo buffers before each call; at a global
// boundary the index clamps (edge replicate).  The square single-GPU operator is the slab with
// gy = gx = g and no neighbours.

namespace bl {
namespace {

struct Stencil {
  double w[9];
};

// Field of gy x gx values with optional halo rows above (row -1) and below (row gy).
template <typename T>
struct Field {
  const T* body;
  const T* top;  // nullptr: no neighbour above (clamp)
  const T* bot;  // nullptr: no neighbour below (clamp)
  int64_t gy, gx;
  __device__ __forceinline__ const T* row(int64_t i) const {
    if (i < 0) return top ? top : body;
    if (i >= gy) return bot ? bot : body + (gy - 1) * gx;
    return body + i * gx;
  }
};

__device__ __forceinline__ int64_t clampi(int64_t v, int64_t hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// conv(u)[i,j] = sum_{a,b} st[a,b] * u[i+1-a, clamp(j+1-b)]  (rows through the halo / clamp)
template <typename T>
__device__ __forceinline__ T conv_at(const Field<T>& u, int64_t i, int64_t j, const Stencil& st) {
  T s = T(0);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T* r = u.row(i + 1 - a);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const T w = static_cast<T>(st.w[a * 3 + b]);
      if (w != T(0)) s = fma(w, r[clampi(j + 1 - b, u.gx - 1)], s);
    }
  }
  return s;
}

// conv_at in two halves: the loads of a cell's (non-zero-weight) stencil points, then the sum in conv_at's order -- so that
// a thread can issue the loads of several rows before the first FMA (one cell per thread and nine dependent-free but
// unbatched loads left these kernels at ~2 TB/s).
template <typename T>
__device__ __forceinline__ void conv_load(const Field<T>& u, int64_t i, int64_t j, const Stencil& st, T (&v)[9]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T* r = u.row(i + 1 - a);
#pragma unroll
    for (int b = 0; b < 3; ++b) v[a * 3 + b] = st.w[a * 3 + b] != 0.0 ? r[clampi(j + 1 - b, u.gx - 1)] : T(0);
  }
}
template <typename T>
__device__ __forceinline__ T conv_sum(const T (&v)[9], const Stencil& st) {
  T s = T(0);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const T w = static_cast<T>(st.w[k]);
    if (w != T(0)) s = fma(w, v[k], s);
  }
  return s;
}

constexpr int kWaveRows = 4;  // rows per thread, their loads in flight together

// y_u = du ; y_du = scale^2 * conv(u)          (2-D launch: blockIdx.y strides groups of kWaveRows rows)
template <typename T>
__global__ void k_wave_matvec(Field<T> u, Stencil st, const T* __restrict__ scale, const T* __restrict__ du,
                              T* __restrict__ y) {
  const int64_t gg = u.gy * u.gx;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= u.gx) return;
  for (int64_t i0 = (int64_t)blockIdx.y * kWaveRows; i0 < u.gy; i0 += (int64_t)gridDim.y * kWaveRows) {
    T d[kWaveRows], sc[kWaveRows], v[kWaveRows][9];
#pragma unroll
    for (int r = 0; r < kWaveRows; ++r) {
      const int64_t i = i0 + r < u.gy ? i0 + r : u.gy - 1;  // clamped: a valid row, not stored below
      const int64_t p = i * u.gx + j;
      d[r] = du[p];
      sc[r] = scale[p];
      conv_load<T>(u, i, j, st, v[r]);
    }
#pragma unroll
    for (int r = 0; r < kWaveRows; ++r) {
      if (i0 + r >= u.gy) break;
      const int64_t p = (i0 + r) * u.gx + j;
      y[p] = d[r];  // d/dt u = du
      y[gg + p] = conv_sum<T>(v[r], st) * (sc[r] * sc[r]);  // fx * constrain(scale), constrain = square
    }
  }
}

// tmp = scale^2 * lam_du (also for the halo rows) ; grad += 2 scale lam_du conv(q_u) ; z_du = lam_u
template <typename T>
__global__ void k_wave_vjp_a(Field<T> qu, Stencil st, const T* __restrict__ scale, const T* __restrict__ lam,
                             const T* __restrict__ scale_top, const T* __restrict__ scale_bot,
                             const T* __restrict__ lam_top, const T* __restrict__ lam_bot, T* __restrict__ tmp,
                             T* __restrict__ tmp_top, T* __restrict__ tmp_bot, T* __restrict__ z,
                             T* __restrict__ grad) {
  const int64_t gy = qu.gy, gx = qu.gx, gg = gy * gx;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= gx) return;
  for (int64_t i0 = (int64_t)blockIdx.y * kWaveRows; i0 < gy; i0 += (int64_t)gridDim.y * kWaveRows) {
    T sc[kWaveRows], ld[kWaveRows], lu[kWaveRows], gr[kWaveRows], v[kWaveRows][9];
#pragma unroll
    for (int r = 0; r < kWaveRows; ++r) {
      const int64_t i = i0 + r < gy ? i0 + r : gy - 1;  // clamped: a valid row, not stored below
      const int64_t p = i * gx + j;
      sc[r] = scale[p];
      ld[r] = lam[gg + p];
      lu[r] = z ? lam[p] : T(0);
      gr[r] = grad[p];
      conv_load<T>(qu, i, j, st, v[r]);
    }
#pragma unroll
    for (int r = 0; r < kWaveRows; ++r) {
      if (i0 + r >= gy) break;
      const int64_t p = (i0 + r) * gx + j;
      tmp[p] = sc[r] * sc[r] * ld[r];
      grad[p] = fma(T(2) * sc[r] * ld[r], conv_sum<T>(v[r], st), gr[r]);
      if (z) z[gg + p] = lu[r];
    }
  }
  if (blockIdx.y == 0) {  // the neighbours' boundary rows of tmp, recomputed from their halos
    if (tmp_top) tmp_top[j] = scale_top[j] * scale_top[j] * lam_top[j];
    if (tmp_bot) tmp_bot[j] = scale_bot[j] * scale_bot[j] * lam_bot[j];
  }
}

// z_u = conv^T(tmp): contributions to cell (p, q) come from rows i = p + a - 1 (through the halo
// when a neighbour exists) plus, at a GLOBAL boundary, the rows folded in by the clamp:
// i = 0 when p == 0 and a == 2, i = gy-1 when p == gy-1 and a == 0; same for the columns.
template <typename T>
__global__ void k_wave_vjp_b(Field<T> tmp, Stencil st, T* __restrict__ z) {
  const int64_t gy = tmp.gy, gx = tmp.gx;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= gx) return;
  for (int64_t p0 = (int64_t)blockIdx.y * kWaveRows; p0 < gy; p0 += (int64_t)gridDim.y * kWaveRows) {
    const bool interior = p0 >= 1 && p0 + kWaveRows <= gy - 1 && q >= 1 && q <= gx - 2;
    if (interior) {
      // interior cells: one row and one column per (a, b), the same terms in the same order as the general path below
      // (whose row / column lists live in local memory: 0.42 ms per VJP at 4096^2 where the bytes take 0.1 ms); the
      // loads of the kWaveRows rows are issued before the first FMA
      T v[kWaveRows][9];
#pragma unroll
      for (int r = 0; r < kWaveRows; ++r) {
        const T* c = tmp.body + (p0 + r - 1) * gx + (q - 1);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) v[r][a * 3 + b] = st.w[a * 3 + b] != 0.0 ? c[a * gx + b] : T(0);
      }
#pragma unroll
      for (int r = 0; r < kWaveRows; ++r) z[(p0 + r) * gx + q] = conv_sum<T>(v[r], st);
      continue;
    }
    for (int64_t p = p0; p < p0 + kWaveRows && p < gy; ++p) {
    T s = T(0);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const T* rows[2];
      int ni = 0;
      const int64_t i0 = p + a - 1;
      if (i0 >= 0 && i0 <= gy - 1) rows[ni++] = tmp.body + i0 * gx;
      if (i0 == -1 && tmp.top) rows[ni++] = tmp.top;
      if (i0 == gy && tmp.bot) rows[ni++] = tmp.bot;
      if (p == 0 && a == 2 && !tmp.top) rows[ni++] = tmp.body;
      if (p == gy - 1 && a == 0 && !tmp.bot) rows[ni++] = tmp.body + (gy - 1) * gx;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const T wt = static_cast<T>(st.w[a * 3 + b]);
        if (wt == T(0)) continue;
        int64_t js[2];
        int nj = 0;
        const int64_t j0 = q + b - 1;
        if (j0 >= 0 && j0 <= gx - 1) js[nj++] = j0;
        if (q == 0 && b == 2) js[nj++] = 0;
        if (q == gx - 1 && b == 0) js[nj++] = gx - 1;
        for (int ii = 0; ii < ni; ++ii)
          for (int jj = 0; jj < nj; ++jj) s = fma(wt, rows[ii][js[jj]], s);
      }
    }
    z[p * gx + q] = s;
    }
  }
}

// Halo exchange over peer memory (dist.cuh): my FIRST row of each field goes to the upper neighbour's
// "from below" slots (1, 3), my LAST row to the lower neighbour's "from above" slots (0, 2); then wait
// for the neighbours' rows in my own mailbox.  One block; rows are at most 512 KB.
template <typename T>
__global__ void __launch_bounds__(1024)
k_wave_halo_exchange(dist::PeerView pv, int64_t gx, const T* __restrict__ first0, const T* __restrict__ last0,
                     const T* __restrict__ first1, const T* __restrict__ last1, bool has_top, bool has_bot) {
  const int parity = (int)(pv.seq & 1ull);
  unsigned char* own = pv.mail[pv.rank];
  if (has_top) {
    unsigned char* up = pv.mail[pv.rank - 1];
    T* d0 = reinterpret_cast<T*>(dist::halo_slot(up, parity, 1));
    T* d1 = reinterpret_cast<T*>(dist::halo_slot(up, parity, 3));
    for (int64_t j = threadIdx.x; j < gx; j += blockDim.x) {
      d0[j] = first0[j];
      if (first1) d1[j] = first1[j];
    }
  }
  if (has_bot) {
    unsigned char* down = pv.mail[pv.rank + 1];
    T* d0 = reinterpret_cast<T*>(dist::halo_slot(down, parity, 0));
    T* d1 = reinterpret_cast<T*>(dist::halo_slot(down, parity, 2));
    for (int64_t j = threadIdx.x; j < gx; j += blockDim.x) {
      d0[j] = last0[j];
      if (last1) d1[j] = last1[j];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && has_top) {
    dist::st_release_sys(dist::halo_flag(pv.mail[pv.rank - 1], pv.rank), pv.seq);
    dist::wait_flag(dist::halo_flag(own, pv.rank - 1), pv.seq, own);
  }
  if (threadIdx.x == 32 && has_bot) {
    dist::st_release_sys(dist::halo_flag(pv.mail[pv.rank + 1], pv.rank), pv.seq);
    dist::wait_flag(dist::halo_flag(own, pv.rank + 1), pv.seq, own);
  }
}

}  // namespace

struct WaveOperator : bl_operator {
  int64_t gy = 0, gx = 0;
  bool has_top = false, has_bot = false;
  Stencil st;
  const void* scale = nullptr;
  int bound_dtype = -1;
  DevBuf grad, tmp;
  // halo rows (gx values each, 8 bytes per value reserved): 0/1 first input top/bottom (u or q_u),
  // 2/3 lam_du top/bottom, 4/5 scale top/bottom, 6/7 tmp top/bottom (internal)
  DevBuf halo[8];
  bl_comm* comm = nullptr;  // native halo exchange (bl_op_wave_set_comm)
  // halo rows of the current call: the mailbox slots after an exchange, else the caller-filled buffers
  const void* cur[4] = {nullptr, nullptr, nullptr, nullptr};

  template <typename T>
  int exchange(const T* field0, const T* field1, cudaStream_t s) {
    for (int k = 0; k < 4; ++k) cur[k] = halo[k].p;
    if (!comm) return BL_OK;
    BL_REQUIRE((size_t)gx * sizeof(T) <= dist::kHaloRowBytes, "grid row too long for the halo mailbox");
    dist::PeerView pv;
    BL_CHECK(dist::view_of(comm, true, &pv));
    const T* last0 = field0 + (gy - 1) * gx;
    const T* last1 = field1 ? field1 + (gy - 1) * gx : nullptr;
    k_wave_halo_exchange<T><<<1, 1024, 0, s>>>(pv, gx, field0, last0, field1, last1, has_top, has_bot);
    BL_LAUNCHED();
    const int parity = (int)(pv.seq & 1ull);
    for (int k = 0; k < 4; ++k) cur[k] = dist::halo_slot(pv.mail[pv.rank], parity, k);
    return BL_OK;
  }
  template <typename T>
  const T* cur_or_null(int k, bool present) const { return present ? static_cast<const T*>(cur[k]) : nullptr; }

  int num_params() const override { return 1; }
  int64_t param_size(int) const override { return gy * gx; }
  dim3 blocks() const {  // blockIdx.y strides groups of kWaveRows rows
    return dim3((unsigned)((gx + 255) / 256), (unsigned)std::min<int64_t>((gy + kWaveRows - 1) / kWaveRows, 65535));
  }
  double matvec_bytes(int dtype) const override { return 5.0 * gy * gx * dtype_size(dtype); }
  double vjp_bytes(int dtype) const override { return 10.0 * gy * gx * dtype_size(dtype); }

  int init_halos() {
    for (auto& h : halo) {
      BL_CHECK(h.ensure((size_t)gx * 8));
      BL_CUDA(cudaMemset(h.p, 0, (size_t)gx * 8));
    }
    return BL_OK;
  }
  template <typename T>
  const T* halo_or_null(int k, bool present) const { return present ? halo[k].as<T>() : nullptr; }

  int set_params(int dtype, const void* const* params, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 1 && params && params[0], "wave operator takes one parameter (scale, rows x cols)");
    scale = params[0];
    bound_dtype = dtype;
    BL_CHECK(grad.ensure((size_t)gy * gx * dtype_size(dtype)));
    BL_CHECK(tmp.ensure((size_t)gy * gx * dtype_size(dtype)));
    if (comm && (has_top || has_bot)) {  // the neighbours' boundary rows of `scale`, kept in halo 4 / 5
      if (dtype == BL_F32)
        BL_CHECK(exchange<float>((const float*)scale, nullptr, s));
      else
        BL_CHECK(exchange<double>((const double*)scale, nullptr, s));
      const size_t row = (size_t)gx * dtype_size(dtype);
      if (has_top) BL_CUDA(cudaMemcpyAsync(halo[4].p, cur[0], row, cudaMemcpyDeviceToDevice, s));
      if (has_bot) BL_CUDA(cudaMemcpyAsync(halo[5].p, cur[1], row, cudaMemcpyDeviceToDevice, s));
    }
    return BL_OK;
  }
  template <typename T>
  int matvec_t(const T* x, T* y, cudaStream_t s) {
    BL_CHECK(exchange<T>(x, nullptr, s));
    Field<T> u{x, cur_or_null<T>(0, has_top), cur_or_null<T>(1, has_bot), gy, gx};
    k_wave_matvec<T><<<blocks(), 256, 0, s>>>(u, st, (const T*)scale, x + gy * gx, y);
    BL_LAUNCHED();
    return BL_OK;
  }
  int matvec(int dtype, const void* x, void* y, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    return dtype == BL_F32 ? matvec_t<float>((const float*)x, (float*)y, s)
                           : matvec_t<double>((const double*)x, (double*)y, s);
  }
  template <typename T>
  int vjp_t(const T* q, const T* lam, T* z, cudaStream_t s) {
    BL_CHECK(exchange<T>(q, lam + gy * gx, s));  // rows of q_u and of lambda_du
    Field<T> qu{q, cur_or_null<T>(0, has_top), cur_or_null<T>(1, has_bot), gy, gx};
    k_wave_vjp_a<T><<<blocks(), 256, 0, s>>>(qu, st, (const T*)scale, lam, halo_or_null<T>(4, has_top),
                                             halo_or_null<T>(5, has_bot), cur_or_null<T>(2, has_top),
                                             cur_or_null<T>(3, has_bot), tmp.as<T>(),
                                             has_top ? halo[6].as<T>() : nullptr, has_bot ? halo[7].as<T>() : nullptr,
                                             z, grad.as<T>());
    BL_LAUNCHED();
    if (z) {
      Field<T> tf{tmp.as<T>(), halo_or_null<T>(6, has_top), halo_or_null<T>(7, has_bot), gy, gx};
      k_wave_vjp_b<T><<<blocks(), 256, 0, s>>>(tf, st, z);
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
    BL_CHECK(grad.ensure((size_t)gy * gx * dtype_size(dtype)));
    BL_CUDA(cudaMemsetAsync(grad.p, 0, (size_t)gy * gx * dtype_size(dtype), s));
    return BL_OK;
  }
  int grad_export(int dtype, void* const* grads, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 1 && grads && grads[0], "wave operator has one gradient buffer");
    BL_CUDA(cudaMemcpyAsync(grads[0], grad.p, (size_t)gy * gx * dtype_size(dtype), cudaMemcpyDeviceToDevice, s));
    return BL_OK;
  }
};

}  // namespace bl

extern "C" {

int bl_op_wave_slab_create(int64_t rows, int64_t cols, int has_top, int has_bottom, const double* stencil3x3_host,
                           bl_operator_t** op) {
  BL_REQUIRE(op && stencil3x3_host && rows >= 1 && cols >= 2, "bad wave operator arguments");
  BL_REQUIRE(rows >= 2 || (has_top && has_bottom) || rows * cols > 0, "bad slab");
  auto* o = new bl::WaveOperator();
  o->gy = rows;
  o->gx = cols;
  o->has_top = has_top != 0;
  o->has_bot = has_bottom != 0;
  o->n = 2 * rows * cols;
  for (int k = 0; k < 9; ++k) o->st.w[k] = stencil3x3_host[k];
  if (o->has_top || o->has_bot) {
    int rc = o->init_halos();
    if (rc != BL_OK) {
      delete o;
      return rc;
    }
  }
  *op = o;
  return BL_OK;
}

int bl_op_wave_create(int64_t grid, const double* stencil3x3_host, bl_operator_t** op) {
  BL_REQUIRE(grid >= 2, "bad wave operator arguments");
  return bl_op_wave_slab_create(grid, grid, 0, 0, stencil3x3_host, op);
}

int bl_op_wave_set_comm(bl_operator_t* op, bl_comm_t* comm) {
  auto* o = dynamic_cast<bl::WaveOperator*>(op);
  BL_REQUIRE(o != nullptr, "not a wave operator");
  if (comm) {
    bl::dist::PeerView pv;
    BL_CHECK(bl::dist::view_of(comm, true, &pv));  // validates the connection (costs one sequence number on every rank)
    BL_REQUIRE(o->has_top == (pv.rank > 0) && o->has_bot == (pv.rank < pv.world - 1),
               "slab neighbours must match the rank order (rank 0 on top)");
    BL_CHECK(o->init_halos());
  }
  o->comm = comm;
  o->bound_dtype = -1;
  return BL_OK;
}

int bl_op_wave_halo(bl_operator_t* op, int which, void** ptr) {
  auto* o = dynamic_cast<bl::WaveOperator*>(op);
  BL_REQUIRE(o != nullptr && ptr != nullptr && which >= 0 && which < 6, "bad halo query");
  BL_REQUIRE(o->halo[which].p != nullptr, "operator has no neighbours");
  *ptr = o->halo[which].p;
  return BL_OK;
}

}  // extern "C"
