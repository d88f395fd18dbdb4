// Row-sharded sparse operand over peer memory ("one large operator": rows of A split over the GPUs of one
// NVSwitch box, north_star).  Rank r owns rows [r*chunk, (r+1)*chunk) of every vector; a matvec first
// ALL-GATHERS the vector -- every rank pushes its chunk into every peer's window with plain NVLink stores,
// the last block publishes a sequence number per peer and waits for theirs (dist.cuh) -- and then applies
// the local rows of A (a rectangular chunk x (world*chunk) SELL operand).  The adjoint gathers q and
// lambda, applies the local rows of A^T and accumulates the cotangent of the locally owned entries.
// No NCCL launch and no host callback on the path; the dot products of the Krylov loops take the same
// route (k_peer_reduce_epilogue).
#include "dist.cuh"
#include "operators.cuh"

namespace bl {
namespace {

constexpr int kGatherThreads = 256;

// dst_p[rank*count + i] = src[i] for every rank p (own window included); then flags + wait (last block)
template <typename T>
__global__ void __launch_bounds__(kGatherThreads)
k_peer_allgather(dist::PeerView pv, dist::WindowView wv, int slot, const T* __restrict__ src, int64_t count,
                 unsigned int* counter) {
  const int parity = (int)(pv.seq & 1ull);
  using V = typename Vec<T>::type;
  constexpr int VN = Vec<T>::N;
  const int64_t nvec = count / VN;  // count is a multiple of VN (chunk is padded to 32 entries)
  for (int p = 0; p < pv.world; ++p) {
    V* dst = reinterpret_cast<V*>(wv.slot(p, parity, slot)) + (int64_t)pv.rank * nvec;
    const V* s = reinterpret_cast<const V*>(src);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = s[i];
  }
  __threadfence_system();
  if (!last_block_done(counter)) return;
  unsigned char* own = pv.mail[pv.rank];
  const int t = threadIdx.x;
  if (t < pv.world && t != pv.rank) {
    dist::st_release_sys(dist::gather_flag(pv.mail[t], slot, pv.rank), pv.seq);
    dist::wait_flag(dist::gather_flag(own, slot, t), pv.seq, own);
  }
}

}  // namespace

struct ShardedSparseOperator : bl_operator {
  bl_operator* A = nullptr;  // local rows of A   : chunk x width
  bl_operator* B = nullptr;  // local rows of A^T : chunk x width
  bl_comm* comm = nullptr;
  int64_t chunk = 0, width = 0;
  DevBuf counter;
  int bound_dtype = -1;

  int num_params() const override { return 2; }
  int64_t param_size(int i) const override { return i == 0 ? A->param_size(0) : B->param_size(0); }
  double matvec_bytes(int dtype) const override { return A->matvec_bytes(dtype) + 2.0 * width * dtype_size(dtype); }
  double vjp_bytes(int dtype) const override { return A->vjp_bytes(dtype) + B->matvec_bytes(dtype) + 4.0 * width * dtype_size(dtype); }

  int set_params(int dtype, const void* const* params, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 2 && params && params[0] && params[1], "sharded operand takes (values of the local rows of A, of A^T)");
    BL_CHECK(A->set_params(dtype, params, 1, s));
    BL_CHECK(B->set_params(dtype, params + 1, 1, s));
    BL_CHECK(counter.ensure(256));
    BL_CUDA(cudaMemsetAsync(counter.p, 0, 256, s));
    bound_dtype = dtype;
    return BL_OK;
  }

  // all-gather of a local vector into window slot `slot`; returns the gathered vector (own window)
  int gather(int dtype, const void* x_loc, int slot, const void** full, cudaStream_t s) {
    dist::PeerView pv;
    dist::WindowView wv;
    BL_CHECK(dist::gather_view_of(comm, &pv, &wv));
    BL_REQUIRE((size_t)width * dtype_size(dtype) <= wv.slot_bytes, "all-gather window too small");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(64, (chunk + 4 * kGatherThreads - 1) / (4 * kGatherThreads)));
    if (dtype == BL_F32)
      k_peer_allgather<float><<<blocks, kGatherThreads, 0, s>>>(pv, wv, slot, (const float*)x_loc, chunk, counter.as<unsigned int>() + slot);
    else
      k_peer_allgather<double><<<blocks, kGatherThreads, 0, s>>>(pv, wv, slot, (const double*)x_loc, chunk, counter.as<unsigned int>() + slot);
    BL_LAUNCHED();
    *full = wv.slot(pv.rank, (int)(pv.seq & 1ull), slot);
    return BL_OK;
  }

  int matvec(int dtype, const void* x, void* y, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    const void* full = nullptr;
    BL_CHECK(gather(dtype, x, 0, &full, s));
    return A->matvec(dtype, full, y, s);
  }
  int vjp(int dtype, const void* q, const void* lam, void* z, cudaStream_t s) override {
    BL_REQUIRE(dtype == bound_dtype, "set_params must be called with the same dtype first");
    const void *q_full = nullptr, *lam_full = nullptr;
    BL_CHECK(gather(dtype, q, 0, &q_full, s));
    if (z) {
      BL_CHECK(gather(dtype, lam, 1, &lam_full, s));
      BL_CHECK(B->matvec(dtype, lam_full, z, s));  // (A^T lambda) restricted to the local rows
    }
    return A->vjp(dtype, q_full, lam, nullptr, s);  // cotangent of the entries in the local rows of A
  }
  int grad_zero(int dtype, cudaStream_t s) override { return A->grad_zero(dtype, s); }
  int grad_export(int dtype, void* const* grads, int num, cudaStream_t s) override {
    BL_REQUIRE(num == 2 && grads && grads[0] && grads[1], "sharded operand has two gradient buffers");
    BL_CHECK(A->grad_export(dtype, grads, 1, s));
    BL_CUDA(cudaMemsetAsync(grads[1], 0, (size_t)B->param_size(0) * dtype_size(dtype), s));  // A^T's copy carries none
    return BL_OK;
  }
};

}  // namespace bl

extern "C" int bl_op_sharded_sparse_create(bl_operator_t* A_local, bl_operator_t* B_local, bl_comm_t* comm,
                                           int64_t chunk, int64_t width, bl_operator_t** op) {
  BL_REQUIRE(op && A_local && B_local && comm && chunk >= 1 && width >= chunk, "bad sharded operand arguments");
  BL_REQUIRE(A_local->n == chunk && B_local->n == chunk, "local operands must have `chunk` rows");
  BL_REQUIRE(chunk % 32 == 0, "chunk must be a multiple of 32 entries");
  auto* o = new bl::ShardedSparseOperator();
  o->A = A_local;
  o->B = B_local;
  o->comm = comm;
  o->chunk = chunk;
  o->width = width;
  o->n = chunk;
  *op = o;
  return BL_OK;
}
