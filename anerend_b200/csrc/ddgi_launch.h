// ddgi_launch.h — backend shim: CUDA runtime (product) or serial host loops (tests/host_emu debug harness).
#pragma once
#include "ddgi_kernels.cuh"

#include <chrono>
#include <cstdlib>
#include <cstring>

#ifndef DDGI_HOST_EMU
// ---------------------------------------------------------------------------------------------- CUDA
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#define DDGI_API(name) anerend_cuda_ddgi_##name

#define DDGI_KERNEL(kname, body, CtxT)                                                   \
  __global__ void __launch_bounds__(256) kname(uint32_t n, const CtxT ctx) {             \
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;                                  \
    if (i < n) body(i, ctx);                                                             \
  }
#define DDGI_LAUNCH(be, kname, body, n, ctx, block)                                      \
  do { if ((n) > 0) { kname<<<((n) + (block) - 1) / (block), (block), 0, (be).stream>>>((uint32_t)(n), ctx); ++(be).launches; } } while (0)

namespace ddgi {
struct Backend {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[11] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // 6..10: kernels of the RT pass
  cudaEvent_t evCopy = nullptr; bool copyPending = false;
  // asynchronous result read-back: a second stream copies the atlas to the host while the next update traces;
  // whatever writes the atlas next waits for it (fenceDownload)
  cudaStream_t dlStream = nullptr; cudaEvent_t evResult = nullptr, evDownload = nullptr; bool downloadPending = false;
  uint32_t launches = 0;
  cudaError_t lastErr = cudaSuccess;
  bool ok(cudaError_t e) { if (e != cudaSuccess) { lastErr = e; return false; } return true; }
  const char* errStr() const { return cudaGetErrorString(lastErr); }
  bool init(int device) {
    int cnt = 0;
    if (!ok(cudaGetDeviceCount(&cnt)) || cnt == 0) return false;
    if (device >= 0 && !ok(cudaSetDevice(device))) return false;
    if (!ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking))) return false;
    { int dev = 0; cudaMemPool_t pool = nullptr; unsigned long long keep = ~0ull;           // freed scratch stays with the process
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      cudaGetLastError(); }
    for (auto& e : ev) if (!ok(cudaEventCreate(&e))) return false;
    if (!ok(cudaStreamCreateWithFlags(&dlStream, cudaStreamNonBlocking))) return false;
    if (!ok(cudaEventCreateWithFlags(&evResult, cudaEventDisableTiming)) || !ok(cudaEventCreateWithFlags(&evDownload, cudaEventDisableTiming))) return false;
    return ok(cudaEventCreateWithFlags(&evCopy, cudaEventDisableTiming));
  }
  bool d2hBackground(void* dst, const void* src, size_t bytes) {      // ordered after everything queued on `stream` so far
    if (!ok(cudaEventRecord(evResult, stream)) || !ok(cudaStreamWaitEvent(dlStream, evResult, 0))) return false;
    if (!ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dlStream))) return false;
    downloadPending = true;
    return ok(cudaEventRecord(evDownload, dlStream));
  }
  // the same for `blocks` strided pieces (rows x width bytes each; source / destination advance by sstep / dstep per block)
  bool d2hStridedBackground(void* dst, size_t dstep, size_t dpitch, const void* src, size_t sstep, size_t spitch, size_t width, size_t rows, size_t blocks) {
    if (!ok(cudaEventRecord(evResult, stream)) || !ok(cudaStreamWaitEvent(dlStream, evResult, 0))) return false;
    for (size_t b = 0; b < blocks; ++b)
      if (rows && !ok(cudaMemcpy2DAsync((char*)dst + dstep * b, dpitch, (const char*)src + sstep * b, spitch, width, rows, cudaMemcpyDeviceToHost, dlStream))) return false;
    downloadPending = true;
    return ok(cudaEventRecord(evDownload, dlStream));
  }
  bool fenceDownload() { return !downloadPending || ok(cudaStreamWaitEvent(stream, evDownload, 0)); }   // device-side: stream waits, host does not
  bool waitDownload() { if (!downloadPending) return true; downloadPending = false; return ok(cudaEventSynchronize(evDownload)); }
  void destroy() {
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (evCopy) cudaEventDestroy(evCopy); if (evResult) cudaEventDestroy(evResult); if (evDownload) cudaEventDestroy(evDownload);
    if (dlStream) cudaStreamDestroy(dlStream); if (stream) cudaStreamDestroy(stream);
  }
  void waitCopy() { if (copyPending) { cudaEventSynchronize(evCopy); copyPending = false; } }
  void markCopy() { cudaEventRecord(evCopy, stream); copyPending = true; }
  template <class T> bool alloc(T** p, size_t count) { *p = nullptr; return ok(cudaMalloc((void**)p, (count ? count : 1) * sizeof(T))); }
  void free(void* p) { if (p) cudaFree(p); }
  // stream-ordered scratch (build-only buffers): no device synchronisation, memory is kept in the pool between builds
  template <class T> bool allocScratch(T** p, size_t count) { *p = nullptr; return ok(cudaMallocAsync((void**)p, (count ? count : 1) * sizeof(T), stream)); }
  void freeScratch(void* p) { if (p) cudaFreeAsync(p, stream); }
  template <class T> bool allocHostPinned(T** p, size_t count) { return ok(cudaMallocHost((void**)p, (count ? count : 1) * sizeof(T))); }
  void freeHostPinned(void* p) { if (p) cudaFreeHost(p); }
  bool h2d(void* dst, const void* src, size_t bytes) { return bytes == 0 || ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream)); }
  bool d2h(void* dst, const void* src, size_t bytes) {
    return bytes == 0 || (ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream)) && ok(cudaStreamSynchronize(stream)));
  }
  bool d2hAsync(void* dst, const void* src, size_t bytes) { return bytes == 0 || ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream)); }
  bool d2hStrided(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
    return rows == 0 || ok(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, stream));
  }
  bool d2d(void* dst, const void* src, size_t bytes) { return bytes == 0 || ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream)); }
  bool zero(void* p, size_t bytes) { return bytes == 0 || ok(cudaMemsetAsync(p, 0, bytes, stream)); }
  bool fill(void* p, int byte, size_t bytes) { return bytes == 0 || ok(cudaMemsetAsync(p, byte, bytes, stream)); }
  bool sync() { return ok(cudaStreamSynchronize(stream)); }
  bool check() { return ok(cudaGetLastError()); }
  // timing events (anerend_cuda_ddgi_set_timers): 0 none, 1 around the passes (what FrameGraphBuilder's timers measure), 2 also around
  // every kernel of the RT pass (events 6..10).  A timing event between two launches is not free: eleven of them cost an update of C3
  // 26 us, and their time-stamp writes queue up behind a read-back of the atlas that is in flight (0.96 instead of 0.87 ms end to end).
  int timerLevel = 1;
  void mark(int i) { if (timerLevel >= (i >= 6 ? 2 : 1)) cudaEventRecord(ev[i], stream); }
  float elapsed(int a, int b) { float ms = 0.0f; if (cudaEventElapsedTime(&ms, ev[a], ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.0f; } return ms; }
  // *tmpOut: the sort's scratch, to be freed by the caller once the stream has drained (no synchronisation here)
  bool sortPairs(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, uint32_t n, void** tmpOut) {
    // 30-bit Morton keys.  CUB's DeviceRadixSort is the onesweep (single-pass, decoupled look-back) radix sort.
    size_t bytes = 0; *tmpOut = nullptr;
    if (!ok(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, 30, stream))) return false;
    if (!ok(cudaMallocAsync(tmpOut, bytes ? bytes : 1, stream))) return false;
    launches += 4;
    return ok(cub::DeviceRadixSort::SortPairs(*tmpOut, bytes, kin, kout, vin, vout, (int)n, 0, 30, stream));
  }
  // exclusive prefix sum (CUB decoupled look-back scan); `tmp` is caller-provided scratch of scanTempBytes(n)
  size_t scanTempBytes(uint32_t n) {
    size_t bytes = 0; cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, stream);
    return bytes ? bytes : 1;
  }
  bool exclusiveSum(void* tmp, size_t tmpBytes, const uint32_t* in, uint32_t* out, uint32_t n) {
    launches += 2;
    return ok(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, in, out, (int)n, stream));
  }
};
}  // namespace ddgi

#else
// ------------------------------------------------------------------------------------------ host emu
#include <algorithm>
#include <numeric>
#include <vector>
#define DDGI_API(name) anerend_emu_ddgi_##name
#define DDGI_KERNEL(kname, body, CtxT)
#define DDGI_LAUNCH(be, kname, body, n, ctx, block) \
  do { for (uint32_t i_ = 0; i_ < (uint32_t)(n); ++i_) body(i_, ctx); ++(be).launches; } while (0)

namespace ddgi {
struct Backend {
  void* stream = nullptr;
  uint32_t launches = 0;
  std::chrono::steady_clock::time_point tp[11];
  const char* errStr() const { return "host emulation"; }
  bool init(int) { return true; }
  void destroy() {}
  void waitCopy() {}
  void markCopy() {}
  bool d2hBackground(void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); return true; }
  bool d2hStridedBackground(void* d, size_t dstep, size_t dp, const void* s, size_t sstep, size_t sp, size_t w, size_t rows, size_t blocks) {
    for (size_t b = 0; b < blocks; ++b) d2hStrided((char*)d + dstep * b, dp, (const char*)s + sstep * b, sp, w, rows);
    return true;
  }
  bool fenceDownload() { return true; }
  bool waitDownload() { return true; }
  template <class T> bool alloc(T** p, size_t count) { *p = (T*)calloc(count ? count : 1, sizeof(T)); return *p != nullptr; }
  void free(void* p) { ::free(p); }
  template <class T> bool allocHostPinned(T** p, size_t count) { return alloc(p, count); }
  void freeHostPinned(void* p) { ::free(p); }
  template <class T> bool allocScratch(T** p, size_t count) { return alloc(p, count); }
  void freeScratch(void* p) { ::free(p); }
  bool h2d(void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); return true; }
  bool d2h(void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); return true; }
  bool d2hAsync(void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); return true; }
  bool d2hStrided(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t rows) {
    for (size_t r = 0; r < rows; ++r) memcpy((char*)d + dp * r, (const char*)s + sp * r, w);
    return true;
  }
  bool d2d(void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); return true; }
  bool zero(void* p, size_t b) { if (b) memset(p, 0, b); return true; }
  bool fill(void* p, int byte, size_t b) { if (b) memset(p, byte, b); return true; }
  bool sync() { return true; }
  bool check() { return true; }
  void mark(int i) { tp[i] = std::chrono::steady_clock::now(); }
  int timerLevel = 1;
  float elapsed(int a, int b) { return std::chrono::duration<float, std::milli>(tp[b] - tp[a]).count(); }
  bool sortPairs(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, uint32_t n, void** tmpOut) {
    *tmpOut = nullptr;
    std::vector<uint32_t> idx(n); std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
    for (uint32_t i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return true;
  }
  size_t scanTempBytes(uint32_t) { return 1; }
  bool exclusiveSum(void*, size_t, const uint32_t* in, uint32_t* out, uint32_t n) {
    uint32_t acc = 0; for (uint32_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = acc; acc += v; }
    return true;
  }
};
}  // namespace ddgi
#endif
