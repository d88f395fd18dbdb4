// Device runtime wrappers of the C ABI (memory, streams, events) and error plumbing.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace bl {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

namespace {
struct ProfRecord {
  int cls;
  double bytes;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRecord> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_enabled() { return g_prof_on; }
void prof_start(int cls, double bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRecord r{cls, bytes, prof_event(), prof_event()};
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void prof_stop(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, s);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
  }
  return cached;
}

}  // namespace bl

extern "C" {

const char* bl_last_error(void) { return bl::t_error.c_str(); }
const char* bl_version(void) { return "b200-lanczos 0.1 (sm_100a)"; }

int bl_device_count(int* count) {
  BL_REQUIRE(count != nullptr, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return BL_OK;
}
int bl_set_device(int device) {
  BL_CUDA(cudaSetDevice(device));
  return BL_OK;
}
int bl_get_device(int* device) {
  BL_REQUIRE(device != nullptr, "device is NULL");
  BL_CUDA(cudaGetDevice(device));
  return BL_OK;
}
int bl_device_sm_count(int* count) {
  BL_REQUIRE(count != nullptr, "count is NULL");
  *count = bl::sm_count();
  return BL_OK;
}
int bl_malloc(void** ptr, size_t bytes) {
  BL_REQUIRE(ptr != nullptr, "ptr is NULL");
  cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    bl::set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? BL_ENOMEM : BL_ECUDA;
  }
  return BL_OK;
}
int bl_free(void* ptr) {
  if (ptr) BL_CUDA(cudaFree(ptr));
  return BL_OK;
}
int bl_host_alloc(void** ptr, size_t bytes) {
  BL_REQUIRE(ptr != nullptr, "ptr is NULL");
  BL_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return BL_OK;
}
int bl_host_free(void* ptr) {
  if (ptr) BL_CUDA(cudaFreeHost(ptr));
  return BL_OK;
}
int bl_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream) {
  if (bytes) BL_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, bl::as_stream(stream)));
  return BL_OK;
}
int bl_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream) {
  if (bytes) BL_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, bl::as_stream(stream)));
  return BL_OK;
}
int bl_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes) BL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, bl::as_stream(stream)));
  return BL_OK;
}
int bl_memset(void* dst, int value, size_t bytes, void* stream) {
  if (bytes) BL_CUDA(cudaMemsetAsync(dst, value, bytes, bl::as_stream(stream)));
  return BL_OK;
}
int bl_stream_create(void** stream) {
  BL_REQUIRE(stream != nullptr, "stream is NULL");
  cudaStream_t s;
  BL_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return BL_OK;
}
int bl_stream_destroy(void* stream) {
  if (stream) BL_CUDA(cudaStreamDestroy(bl::as_stream(stream)));
  return BL_OK;
}
int bl_stream_sync(void* stream) {
  BL_CUDA(cudaStreamSynchronize(bl::as_stream(stream)));
  return BL_OK;
}
int bl_device_sync(void) {
  BL_CUDA(cudaDeviceSynchronize());
  return BL_OK;
}
int bl_event_create(void** event) {
  BL_REQUIRE(event != nullptr, "event is NULL");
  cudaEvent_t e;
  BL_CUDA(cudaEventCreate(&e));
  *event = e;
  return BL_OK;
}
int bl_event_destroy(void* event) {
  if (event) BL_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return BL_OK;
}
int bl_event_record(void* event, void* stream) {
  BL_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), bl::as_stream(stream)));
  return BL_OK;
}
int bl_stream_wait_event(void* stream, void* event) {
  BL_CUDA(cudaStreamWaitEvent(bl::as_stream(stream), static_cast<cudaEvent_t>(event), 0));
  return BL_OK;
}
int bl_event_sync(void* event) {
  BL_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return BL_OK;
}
int bl_event_elapsed_ms(void* start, void* stop, float* ms) {
  BL_REQUIRE(ms != nullptr, "ms is NULL");
  BL_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return BL_OK;
}
int bl_profile_begin(void) {
  std::lock_guard<std::mutex> lk(bl::g_prof_mu);
  for (auto& r : bl::g_prof) {
    bl::g_prof_pool.push_back(r.a);
    bl::g_prof_pool.push_back(r.b);
  }
  bl::g_prof.clear();
  bl::g_prof_on = true;
  return BL_OK;
}
int bl_profile_end(uint64_t* counts, double* ms, double* bytes) {
  BL_REQUIRE(counts && ms && bytes, "NULL argument");
  BL_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(bl::g_prof_mu);
  bl::g_prof_on = false;
  for (int c = 0; c < BL_PROF_NCLASS; ++c) counts[c] = 0, ms[c] = 0.0, bytes[c] = 0.0;
  for (auto& r : bl::g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      counts[r.cls] += 1;
      ms[r.cls] += t;
      bytes[r.cls] += r.bytes;
    }
    bl::g_prof_pool.push_back(r.a);
    bl::g_prof_pool.push_back(r.b);
  }
  bl::g_prof.clear();
  return BL_OK;
}
int bl_launch_count(uint64_t* count) {
  BL_REQUIRE(count != nullptr, "count is NULL");
  *count = bl::g_launches.load();
  return BL_OK;
}

}  // extern "C"
