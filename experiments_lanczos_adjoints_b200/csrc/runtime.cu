// Device runtime wrappers of the C ABI (memory, streams, events) and error plumbing.
#include <mutex>

#include "common.cuh"

namespace bl {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

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
int bl_event_sync(void* event) {
  BL_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return BL_OK;
}
int bl_event_elapsed_ms(void* start, void* stop, float* ms) {
  BL_REQUIRE(ms != nullptr, "ms is NULL");
  BL_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return BL_OK;
}
int bl_launch_count(uint64_t* count) {
  BL_REQUIRE(count != nullptr, "count is NULL");
  *count = bl::g_launches.load();
  return BL_OK;
}

}  // extern "C"
