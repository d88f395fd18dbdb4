// Peer-memory communicator: mailbox allocation, CUDA-IPC mapping, activation (see dist.cuh).
#include "dist.cuh"

struct bl_comm {
  int rank = 0, world = 1, device = 0;
  unsigned char* mail[bl::dist::kMaxRanks] = {};
  bool ipc[bl::dist::kMaxRanks] = {};
  unsigned long long red_seq = 0, halo_seq = 0, gather_seq = 0;
  bool header_written = false;
  size_t slot_bytes = 0;  // all-gather window: [2][2][slot_bytes] per rank
  unsigned char* win[bl::dist::kMaxRanks] = {};
  bool win_ipc[bl::dist::kMaxRanks] = {};
};

namespace bl {
namespace dist {

static thread_local bl_comm* t_comm = nullptr;

bool active() { return t_comm != nullptr; }
int world() { return t_comm ? t_comm->world : 1; }

int view_of(bl_comm* c, bool halo, PeerView* pv) {
  BL_REQUIRE(c != nullptr, "no communicator");
  for (int p = 0; p < c->world; ++p) BL_REQUIRE(c->mail[p] != nullptr, "communicator is not connected to every rank");
  pv->rank = c->rank;
  pv->world = c->world;
  pv->seq = halo ? ++c->halo_seq : ++c->red_seq;
  for (int p = 0; p < kMaxRanks; ++p) pv->mail[p] = c->mail[p];
  return BL_OK;
}
int gather_view_of(bl_comm* c, PeerView* pv, WindowView* wv) {
  BL_REQUIRE(c != nullptr && c->slot_bytes > 0, "communicator has no all-gather window (bl_dist_comm_window_create)");
  for (int p = 0; p < c->world; ++p) BL_REQUIRE(c->mail[p] && c->win[p], "communicator / window is not connected to every rank");
  pv->rank = c->rank;
  pv->world = c->world;
  pv->seq = ++c->gather_seq;
  wv->slot_bytes = c->slot_bytes;
  for (int p = 0; p < kMaxRanks; ++p) pv->mail[p] = c->mail[p], wv->win[p] = c->win[p];
  return BL_OK;
}
int next_reduce(PeerView* pv) { return view_of(t_comm, false, pv); }
int next_halo(PeerView* pv) { return view_of(t_comm, true, pv); }

}  // namespace dist
}  // namespace bl

using namespace bl;

extern "C" {

int bl_dist_comm_create(int rank, int world, bl_comm_t** comm) {
  BL_REQUIRE(comm != nullptr && world >= 1 && world <= dist::kMaxRanks && rank >= 0 && rank < world,
             "bad communicator arguments (at most 8 ranks)");
  auto* c = new bl_comm();
  c->rank = rank;
  c->world = world;
  if (cudaGetDevice(&c->device) != cudaSuccess || cudaMalloc(&c->mail[rank], dist::kMailboxBytes) != cudaSuccess ||
      cudaMemset(c->mail[rank], 0, dist::kMailboxBytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("mailbox allocation failed");
    delete c;
    return BL_ENOMEM;
  }
  *comm = c;
  return BL_OK;
}

int bl_dist_comm_local(bl_comm_t* comm, void** mailbox, void* ipc_handle_64) {
  BL_REQUIRE(comm != nullptr, "no communicator");
  if (mailbox) *mailbox = comm->mail[comm->rank];
  if (ipc_handle_64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    BL_CUDA(cudaIpcGetMemHandle(&h, comm->mail[comm->rank]));
    std::memcpy(ipc_handle_64, &h, 64);
  }
  return BL_OK;
}

int bl_dist_comm_connect_ipc(bl_comm_t* comm, const void* handles) {
  BL_REQUIRE(comm != nullptr && handles != nullptr, "bad arguments");
  for (int p = 0; p < comm->world; ++p) {
    if (p == comm->rank || comm->mail[p]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + 64 * p, 64);
    void* ptr = nullptr;
    BL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    comm->mail[p] = static_cast<unsigned char*>(ptr);
    comm->ipc[p] = true;
  }
  return BL_OK;
}

int bl_dist_comm_connect_ptrs(bl_comm_t* comm, void* const* mailboxes) {
  BL_REQUIRE(comm != nullptr && mailboxes != nullptr, "bad arguments");
  for (int p = 0; p < comm->world; ++p) {
    if (p == comm->rank) continue;
    BL_REQUIRE(mailboxes[p] != nullptr, "missing mailbox pointer");
    comm->mail[p] = static_cast<unsigned char*>(mailboxes[p]);
  }
  return BL_OK;
}

int bl_dist_comm_window_create(bl_comm_t* comm, size_t slot_bytes, void* ipc_handle_64) {
  BL_REQUIRE(comm != nullptr && slot_bytes > 0 && comm->slot_bytes == 0, "bad window arguments (one window per communicator)");
  slot_bytes = align_up(slot_bytes, 256);
  void* p = nullptr;
  if (cudaMalloc(&p, 4 * slot_bytes) != cudaSuccess || cudaMemset(p, 0, 4 * slot_bytes) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    set_error("window allocation failed");
    return BL_ENOMEM;
  }
  comm->win[comm->rank] = static_cast<unsigned char*>(p);
  comm->slot_bytes = slot_bytes;
  if (ipc_handle_64) {
    cudaIpcMemHandle_t h;
    BL_CUDA(cudaIpcGetMemHandle(&h, p));
    std::memcpy(ipc_handle_64, &h, 64);
  }
  return BL_OK;
}

int bl_dist_comm_window_local(bl_comm_t* comm, void** window) {
  BL_REQUIRE(comm != nullptr && window != nullptr && comm->slot_bytes > 0, "no window");
  *window = comm->win[comm->rank];
  return BL_OK;
}

int bl_dist_comm_window_connect_ipc(bl_comm_t* comm, const void* handles) {
  BL_REQUIRE(comm != nullptr && handles != nullptr && comm->slot_bytes > 0, "bad arguments");
  for (int p = 0; p < comm->world; ++p) {
    if (p == comm->rank || comm->win[p]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + 64 * p, 64);
    void* ptr = nullptr;
    BL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    comm->win[p] = static_cast<unsigned char*>(ptr);
    comm->win_ipc[p] = true;
  }
  return BL_OK;
}

int bl_dist_comm_window_connect_ptrs(bl_comm_t* comm, void* const* windows) {
  BL_REQUIRE(comm != nullptr && windows != nullptr && comm->slot_bytes > 0, "bad arguments");
  for (int p = 0; p < comm->world; ++p) {
    if (p == comm->rank) continue;
    BL_REQUIRE(windows[p] != nullptr, "missing window pointer");
    comm->win[p] = static_cast<unsigned char*>(windows[p]);
  }
  return BL_OK;
}

int bl_dist_comm_activate(bl_comm_t* comm) {
  if (comm && !comm->header_written) {  // publish the peer table in the own mailbox (kernels rebuild their PeerView from it)
    dist::MailHeader h{};
    h.rank = comm->rank;
    h.world = comm->world;
    for (int p = 0; p < comm->world; ++p) {
      BL_REQUIRE(comm->mail[p] != nullptr, "communicator is not connected to every rank");
      h.mail[p] = comm->mail[p];
    }
    BL_CUDA(cudaMemcpy(comm->mail[comm->rank] + dist::kHeaderOff, &h, sizeof(h), cudaMemcpyHostToDevice));
    comm->header_written = true;
  }
  dist::t_comm = comm;
  return BL_OK;
}

int bl_dist_comm_error(bl_comm_t* comm, int* timed_out) {
  BL_REQUIRE(comm != nullptr && timed_out != nullptr, "bad arguments");
  unsigned long long flag = 0;
  BL_CUDA(cudaMemcpy(&flag, comm->mail[comm->rank] + 64 * 8, 8, cudaMemcpyDeviceToHost));
  *timed_out = flag != 0;
  return BL_OK;
}

int bl_dist_comm_destroy(bl_comm_t* comm) {
  if (!comm) return BL_OK;
  if (dist::t_comm == comm) dist::t_comm = nullptr;
  for (int p = 0; p < comm->world; ++p) {
    if (!comm->mail[p]) continue;
    if (p == comm->rank)
      cudaFree(comm->mail[p]);
    else if (comm->ipc[p])
      cudaIpcCloseMemHandle(comm->mail[p]);
  }
  for (int p = 0; p < comm->world; ++p) {
    if (!comm->win[p]) continue;
    if (p == comm->rank)
      cudaFree(comm->win[p]);
    else if (comm->win_ipc[p])
      cudaIpcCloseMemHandle(comm->win[p]);
  }
  delete comm;
  return BL_OK;
}

}  // extern "C"
