// Execution backends for the phasing pipeline.
//
// The pipeline (phz_pipeline.h) is written once against a tiny set of data-parallel primitives:
// for_each (one logical thread per item), exclusive scan, radix sort, memset/copies.  The PRODUCT
// is built with DeviceBackend (CUDA kernels on a stream, CUB for scan/sort) into
// phaser_b200/_phz.so.  HostSimBackend runs the SAME per-thread code serially on the CPU and is
// compiled only into tests/hostsim/_phz_hostsim.so: it is a logic-test double for the build
// container (which has no GPU), never shipped, never loaded by phaser_b200 on its own.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>

#ifdef __CUDACC__
#define PHZ_HD __host__ __device__ __forceinline__
#else
#define PHZ_HD inline
#endif

namespace phz {

// ----------------------------------------------------------------------------- atomics (HD)
#if defined(__CUDA_ARCH__)
PHZ_HD uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
PHZ_HD unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
PHZ_HD uint32_t atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
PHZ_HD unsigned long long atomic_min(unsigned long long* p, unsigned long long v) { return atomicMin(p, v); }
PHZ_HD uint32_t atomic_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
PHZ_HD uint32_t atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
PHZ_HD uint32_t atomic_and(uint32_t* p, uint32_t v) { return atomicAnd(p, v); }
PHZ_HD uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
PHZ_HD unsigned long long atomic_cas(unsigned long long* p, unsigned long long cmp, unsigned long long v) { return atomicCAS(p, cmp, v); }
PHZ_HD uint32_t load_volatile(const uint32_t* p) { return *((const volatile uint32_t*)p); }
PHZ_HD unsigned long long load_volatile(const unsigned long long* p) { return *((const volatile unsigned long long*)p); }
#else
template <class T> inline T atomic_add(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomic_min(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomic_max(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
inline uint32_t atomic_or(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
inline uint32_t atomic_and(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
inline uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
inline unsigned long long atomic_cas(unsigned long long* p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
inline uint32_t load_volatile(const uint32_t* p) { return *p; }
inline unsigned long long load_volatile(const unsigned long long* p) { return *p; }
#endif

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint8_t u8;

// Full-warp helpers for for_each_warp bodies (all 32 lanes of a warp run the body for the same item; the host
// simulation runs it once with one lane).
PHZ_HD u32 warp_ballot(bool p) {
#if defined(__CUDA_ARCH__)
  return __ballot_sync(0xFFFFFFFFu, p);
#else
  return p ? 1u : 0u;
#endif
}
PHZ_HD u32 popc_u32(u32 x) {
#if defined(__CUDA_ARCH__)
  return (u32)__popc(x);
#else
  return (u32)__builtin_popcount(x);
#endif
}

// *addr += 1, combined over whatever lanes of the warp are converged here and target the same address (usable inside
// divergent loops: the set of participants is taken as it is found)
PHZ_HD void converged_inc(u32* addr) {
#if defined(__CUDA_ARCH__)
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, (unsigned long long)addr);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(addr, (u32)__popc(peers));
#else
  *addr += 1;
#endif
}

// counter[key] += 1 for every lane with `active`; lanes of a warp that hit the same key are combined
// into one atomic (neighbouring items of the sorted arrays mostly belong to the same locus).  Must be
// reached by all lanes that are active at the call site.
PHZ_HD void warp_agg_inc(u32* counter, u32 key, bool active) {
#if defined(__CUDA_ARCH__)
  unsigned act = __activemask();
  unsigned peers = __match_any_sync(act, active ? key : 0xFFFFFFFFu);
  if (active && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counter[key], (u32)__popc(peers));
#else
  if (active) counter[key] += 1;
#endif
}

// lower_bound on a sorted int32 array: first index in [lo, hi) with a[idx] >= key
PHZ_HD int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int32_t key) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// last index i in [0, n) with off[i] <= x  (off sorted ascending, off[0] <= x)
PHZ_HD int upper_slot_i64(const int64_t* off, int n, int64_t x) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

struct PhzError : public std::runtime_error {
  explicit PhzError(const std::string& s) : std::runtime_error(s) {}
};

}  // namespace phz

#ifdef __CUDACC__
// =============================================================================== CUDA backend
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/counting_iterator.h>

namespace phz {

template <class T> struct Widen { __host__ __device__ T operator()(uint16_t x) const { return (T)x; } };

#define PHZ_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw PhzError(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +       \
                     __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

template <class F>
__global__ void __launch_bounds__(256) for_each_kernel(int64_t n, F f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}

// one warp per item: f(item, lane, 32)
template <class F>
__global__ void __launch_bounds__(256) for_each_warp_kernel(int64_t n, F f) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w < n) f(w, (int)(threadIdx.x & 31), 32);
}

struct DeviceBackend {
  static constexpr bool kIsDevice = true;
  cudaStream_t stream = nullptr;
  int device = 0;
  u64 launches = 0;          // kernels of OURS launched (for_each + hand-written); CUB passes counted separately
  u64 lib_launches = 0;
  u64 syncs = 0;             // blocking waits of the host on this stream (device counters read back, explicit syncs)
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  // optional per-stage CUDA-event timing on this stream (phz_set_profiling)
  int profiling = 0;          // 1: K1 pass events, 2: + named stage marks
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void mark(int i) {
    if (!profiling) return;
    if (!ev[i]) PHZ_CUDA(cudaEventCreate(&ev[i]));
    PHZ_CUDA(cudaEventRecord(ev[i], stream));
  }
  float elapsed(int a, int b) {
    if (!profiling || !ev[a] || !ev[b]) return -1.f;
    float ms = 0.f;
    PHZ_CUDA(cudaEventSynchronize(ev[b]));
    PHZ_CUDA(cudaEventElapsedTime(&ms, ev[a], ev[b]));
    return ms;
  }
  // named stage marks: stage(name) closes the previous stage and opens `name`
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  void stage(const char* name) {
    if (profiling < 2) return;
    cudaEvent_t e; PHZ_CUDA(cudaEventCreate(&e)); PHZ_CUDA(cudaEventRecord(e, stream));
    marks.emplace_back(name, e);
  }
  std::string stage_report() {
    std::string out;
    if (marks.empty()) return out;
    PHZ_CUDA(cudaEventSynchronize(marks.back().second));
    for (size_t i = 0; i + 1 < marks.size(); ++i) {
      float ms = 0.f; cudaEventElapsedTime(&ms, marks[i].second, marks[i + 1].second);
      out += marks[i].first + "\t" + std::to_string(ms) + "\n";
    }
    for (auto& m : marks) cudaEventDestroy(m.second);
    marks.clear();
    return out;
  }

  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    PHZ_CUDA(cudaMalloc(&p, bytes));
    return p;
  }
  void free(void* p) { if (p) cudaFree(p); }
  // page-locked host memory for transport buffers; plain malloc when no device is usable (ingest-only use of the library)
  static void* host_alloc(size_t bytes, bool* pinned) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) { *pinned = true; return p; }
    cudaGetLastError();
    *pinned = false;
    return std::malloc(bytes);
  }
  static void host_free(void* p, bool pinned) { if (!p) return; if (pinned) cudaFreeHost(p); else std::free(p); }
  void memset0(void* p, size_t bytes) { if (bytes) PHZ_CUDA(cudaMemsetAsync(p, 0, bytes, stream)); }
  void memset_ff(void* p, size_t bytes) { if (bytes) PHZ_CUDA(cudaMemsetAsync(p, 0xFF, bytes, stream)); }
  void h2d(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
  }
  void d2h(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    PHZ_CUDA(cudaStreamSynchronize(stream));
    syncs++;
  }
  void d2h_async(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
  }
  void d2d(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
  }
  // device -> wherever dst lives (host or device; unified addressing decides), no wait
  void copy_out_async(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream));
  }
  // ---- second stream for host->device copies that overlap kernels of the main stream (phz_map_reads_packed):
  // copy_begin() orders the copy stream after everything already queued on the main stream (the staging buffers
  // may still be read by it), h2d_copy() enqueues on the copy stream, copy_fence() makes the main stream wait for
  // the copies enqueued so far.
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  void* up_stage[2] = {nullptr, nullptr};          // page-locked staging of phz_upload
  cudaEvent_t up_ev[2] = {nullptr, nullptr};
  void copy_begin() {
    if (!copy_stream) {
      PHZ_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      for (auto& e : copy_ev) PHZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    PHZ_CUDA(cudaEventRecord(copy_ev[0], stream));
    PHZ_CUDA(cudaStreamWaitEvent(copy_stream, copy_ev[0], 0));
  }
  void h2d_copy(void* dst, const void* src, size_t bytes) {
    if (bytes) PHZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, copy_stream));
  }
  void copy_fence() {
    PHZ_CUDA(cudaEventRecord(copy_ev[1], copy_stream));
    PHZ_CUDA(cudaStreamWaitEvent(stream, copy_ev[1], 0));
  }
  // caller-owned events for copies that outlive the call that enqueued them (prefetch)
  void* new_event() { cudaEvent_t e; PHZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return (void*)e; }
  void free_event(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
  void copy_record(void* ev) { PHZ_CUDA(cudaEventRecord((cudaEvent_t)ev, copy_stream)); }
  void wait_event(void* ev) { PHZ_CUDA(cudaStreamWaitEvent(stream, (cudaEvent_t)ev, 0)); }
  void record_event(void* ev) { PHZ_CUDA(cudaEventRecord((cudaEvent_t)ev, stream)); }       // on the main stream
  void host_wait_event(void* ev) { PHZ_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); }       // any host thread; the stream runs on
  void sync() { PHZ_CUDA(cudaStreamSynchronize(stream)); syncs++; }

  template <class F>
  void for_each(int64_t n, F f) {
    if (n <= 0) return;
    int64_t blocks = (n + 255) / 256;
    for_each_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n, f);
    PHZ_CUDA(cudaGetLastError());
    launches++;
  }

  template <class F>
  void for_each_warp(int64_t n, F f) {
    if (n <= 0) return;
    int64_t blocks = (n * 32 + 255) / 256;
    for_each_warp_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n, f);
    PHZ_CUDA(cudaGetLastError());
    launches++;
  }

  void* tmp(size_t bytes) {
    if (bytes > cub_tmp_bytes) {
      if (cub_tmp) { PHZ_CUDA(cudaStreamSynchronize(stream)); cudaFree(cub_tmp); }
      cub_tmp_bytes = bytes + bytes / 4 + 1024;
      PHZ_CUDA(cudaMalloc(&cub_tmp, cub_tmp_bytes));
    }
    return cub_tmp;
  }

  // out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 slots (in-place NOT allowed).
  void exclusive_scan_u32(const u32* in, u32* out, int64_t n) {
    memset0(out + n, sizeof(u32));
    if (n <= 0) return;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, (int)n, stream));
    lib_launches += 2;
    const u32* in_c = in; u32* out_c = out; int64_t nn = n;
    for_each(1, [=] __device__(int64_t) { out_c[nn] = out_c[nn - 1] + in_c[nn - 1]; });
  }

  // out[i] = sum_{j<i} (u64)in[j] for i in [0, n]; out has n+1 slots
  void exclusive_scan_u16_to_u64(const uint16_t* in, u64* out, int64_t n) {
    memset0(out + n, sizeof(u64));
    if (n <= 0) return;
    thrust::transform_iterator<Widen<u64>, const uint16_t*> it(in, Widen<u64>());
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, it, out, (int)n, stream));
    lib_launches += 2;
    const uint16_t* in_c = in; u64* out_c = out; int64_t nn = n;
    for_each(1, [=] __device__(int64_t) { out_c[nn] = out_c[nn - 1] + in_c[nn - 1]; });
  }
  // out[i] = sum_{j<i} f(j) for i in [0, n]; f(j) in {0, 1, ...} is evaluated on the fly (no flag array in memory)
  template <class F>
  void exclusive_scan_fn_u32(F f, u32* out, int64_t n) {
    memset0(out + n, sizeof(u32));
    if (n <= 0) return;
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), f);
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, it, out, (int)n, stream));
    lib_launches += 2;
    u32* out_c = out; int64_t nn = n;
    for_each(1, [=] __device__(int64_t) { out_c[nn] = out_c[nn - 1] + f(nn - 1); });
  }
  void inclusive_scan_i32_inplace(int32_t* a, int64_t n) {
    if (n <= 0) return;
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, a, a, (int)n, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceScan::InclusiveSum(t, bytes, a, a, (int)n, stream));
    lib_launches += 2;
  }
  void exclusive_scan_u16_to_u32(const uint16_t* in, u32* out, int64_t n) {
    memset0(out + n, sizeof(u32));
    if (n <= 0) return;
    thrust::transform_iterator<Widen<u32>, const uint16_t*> it(in, Widen<u32>());
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, it, out, (int)n, stream));
    lib_launches += 2;
    const uint16_t* in_c = in; u32* out_c = out; int64_t nn = n;
    for_each(1, [=] __device__(int64_t) { out_c[nn] = out_c[nn - 1] + in_c[nn - 1]; });
  }

  template <class V>
  void sort_pairs(u64* keys_in, u64* keys_out, V* vals_in, V* vals_out, int64_t n, int begin_bit, int end_bit) {
    if (n <= 0) return;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, stream));
    lib_launches += 1 + (end_bit - begin_bit + 7) / 8;
  }
  template <class V>
  void sort_pairs32(u32* keys_in, u32* keys_out, V* vals_in, V* vals_out, int64_t n, int begin_bit, int end_bit) {
    if (n <= 0) return;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, stream);
    void* t = tmp(bytes + 16);
    PHZ_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, stream));
    lib_launches += 1 + (end_bit - begin_bit + 7) / 8;
  }
  ~DeviceBackend() {
    if (cub_tmp) cudaFree(cub_tmp);
    for (auto& e : copy_ev) if (e) cudaEventDestroy(e);
    for (auto& e : up_ev) if (e) cudaEventDestroy(e);
    for (auto& p : up_stage) if (p) cudaFreeHost(p);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
};

}  // namespace phz

#else
// =============================================================================== host simulation
namespace phz {

struct HostSimBackend {
  static constexpr bool kIsDevice = false;
  void* stream = nullptr;
  int device = -1;
  u64 launches = 0;
  u64 lib_launches = 0;
  u64 syncs = 0;
  int profiling = 0;
  void mark(int) {}
  float elapsed(int, int) { return -1.f; }
  void stage(const char*) {}
  std::string stage_report() { return std::string(); }

  void* alloc(size_t bytes) { if (bytes == 0) bytes = 16; void* p = std::malloc(bytes); if (!p) throw PhzError("host alloc failed"); return p; }
  void free(void* p) { std::free(p); }
  void memset0(void* p, size_t bytes) { std::memset(p, 0, bytes); }
  void memset_ff(void* p, size_t bytes) { std::memset(p, 0xFF, bytes); }
  void h2d(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }
  void d2h(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); syncs++; }
  void d2h_async(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }
  void d2d(void* dst, const void* src, size_t bytes) { if (bytes) std::memmove(dst, src, bytes); }
  void copy_out_async(void* dst, const void* src, size_t bytes) { if (bytes) std::memmove(dst, src, bytes); }
  void copy_begin() {}
  void h2d_copy(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }
  void copy_fence() {}
  void* new_event() { return (void*)1; }
  void free_event(void*) {}
  void copy_record(void*) {}
  void wait_event(void*) {}
  void record_event(void*) {}
  void host_wait_event(void*) {}
  void sync() {}

  template <class F>
  void for_each(int64_t n, F f) {
    for (int64_t i = 0; i < n; ++i) f(i);
    launches++;
  }
  template <class F>
  void for_each_warp(int64_t n, F f) {
    for (int64_t i = 0; i < n; ++i) f(i, 0, 1);
    launches++;
  }
  void exclusive_scan_u32(const u32* in, u32* out, int64_t n) {
    u32 s = 0;
    for (int64_t i = 0; i < n; ++i) { u32 v = in[i]; out[i] = s; s += v; }
    out[n] = s;
  }
  void exclusive_scan_u16_to_u64(const uint16_t* in, u64* out, int64_t n) {
    u64 s = 0;
    for (int64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
    out[n] = s;
  }
  template <class F>
  void exclusive_scan_fn_u32(F f, u32* out, int64_t n) {
    u32 s = 0;
    for (int64_t i = 0; i < n; ++i) { out[i] = s; s += f(i); }
    out[n] = s;
  }
  void inclusive_scan_i32_inplace(int32_t* a, int64_t n) {
    u32 s = 0;
    for (int64_t i = 0; i < n; ++i) { s += (u32)a[i]; a[i] = (int32_t)s; }
  }
  void exclusive_scan_u16_to_u32(const uint16_t* in, u32* out, int64_t n) {
    u32 s = 0;
    for (int64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
    out[n] = s;
  }
  static void* host_alloc(size_t bytes, bool* pinned) { *pinned = false; return std::malloc(bytes ? bytes : 16); }
  static void host_free(void* p, bool) { std::free(p); }
  template <class K, class V>
  void sort_impl(K* keys_in, K* keys_out, V* vals_in, V* vals_out, int64_t n, int begin_bit, int end_bit) {
    std::vector<int64_t> idx(n);
    for (int64_t i = 0; i < n; ++i) idx[i] = i;
    K mask = (end_bit - begin_bit >= (int)(8 * sizeof(K))) ? ~(K)0 : ((((K)1) << (end_bit - begin_bit)) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
      return ((keys_in[a] >> begin_bit) & mask) < ((keys_in[b] >> begin_bit) & mask);
    });
    for (int64_t i = 0; i < n; ++i) { keys_out[i] = keys_in[idx[i]]; vals_out[i] = vals_in[idx[i]]; }
  }
  template <class V>
  void sort_pairs(u64* ki, u64* ko, V* vi, V* vo, int64_t n, int b, int e) { sort_impl<u64, V>(ki, ko, vi, vo, n, b, e); }
  template <class V>
  void sort_pairs32(u32* ki, u32* ko, V* vi, V* vo, int64_t n, int b, int e) { sort_impl<u32, V>(ki, ko, vi, vo, n, b, e); }
};

}  // namespace phz
#endif

namespace phz {

// Grow-only typed buffer owned by a backend (device memory in the product).
template <class B, class T>
struct Buf {
  B* be = nullptr;
  T* p = nullptr;
  size_t cap = 0;
  Buf() {}
  explicit Buf(B* b) : be(b) {}
  Buf(const Buf&) = delete;
  Buf& operator=(const Buf&) = delete;
  ~Buf() { if (p && be) be->free(p); }
  void bind(B* b) { be = b; }
  // contents are NOT preserved
  T* ensure(size_t n) {
    if (n > cap) {
      if (!be) throw PhzError("internal: buffer used before it was bound to a backend");
      if (p) { be->sync(); be->free(p); }
      cap = n + n / 8 + 64;
      p = (T*)be->alloc(cap * sizeof(T));
    }
    return p;
  }
  // contents preserved
  T* grow(size_t n, size_t used) {
    if (n > cap) {
      if (!be) throw PhzError("internal: buffer used before it was bound to a backend");
      size_t ncap = n + n / 4 + 64;
      T* q = (T*)be->alloc(ncap * sizeof(T));
      if (p && used) be->d2d(q, p, used * sizeof(T));
      if (p) { be->sync(); be->free(p); }
      p = q; cap = ncap;
    }
    return p;
  }
  void release() { if (p) { be->sync(); be->free(p); p = nullptr; cap = 0; } }
};

}  // namespace phz
