// fake_cudart.cpp — a single-threaded model of the part of the CUDA runtime that pl-svo_b200/csrc/plsvo_abi.cu uses.
// TEST INFRASTRUCTURE ONLY (see fake_cuda.h).
//
// What is modelled, because the host pipeline's correctness depends on it:
//   * streams are FIFO queues of operations; nothing runs at enqueue time in the default ("lazy") mode — an operation runs
//     only when a synchronising call (cudaStreamSynchronize, cudaEventSynchronize) forces its stream forward, and then the
//     other streams advance only as far as events and the arrival gate require.  That is the most adversarial legal
//     schedule for "was this dependency expressed?": a copy the kernel needs but never waits for has simply not happened.
//     PLSVO_FAKE_CUDA=eager runs every operation as early as its dependencies allow (the other extreme: a copy that may
//     overtake work still using its destination does so).
//   * events carry the generation of their last cudaEventRecord; cudaStreamWaitEvent captures the generation current at
//     the time of the call, as CUDA does.
//   * "device" and pinned memory are heap blocks filled with 0xCD at allocation, and every copy / memset / model kernel is
//     bounds-checked against the block it touches: an undersized buffer or a frame that never arrived shows up as a
//     recorded error or as poison in a digest.
//   * host->device copies read their source when they RUN (page-locked semantics for every source): a call that returns
//     while such a copy is pending is reported by fake_cuda_pending_host_reads().
#include "fake_cuda.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <map>
#include <vector>

namespace {

struct Op {
  fakecuda::OpFn run;
  bool reads_host;
};

struct Stream {
  std::deque<Op> q;
  bool busy = false;  // an operation of this stream is running (it may be advancing other streams)
  bool alive = true;
};

struct Event {
  uint64_t gen_enqueued = 0, gen_done = 0;
};

std::vector<Stream*> g_streams;
Stream g_default;  // the legacy default stream (stream handle 0)
std::map<uintptr_t, size_t> g_blocks;  // base -> size of every live device / pinned block
std::string g_errors;
int g_eager = -1;
uint64_t g_ops_run = 0, g_bytes_h2d = 0;

bool eager() {
  if (g_eager < 0) {
    const char* e = getenv("PLSVO_FAKE_CUDA");
    g_eager = (e && strcmp(e, "eager") == 0) ? 1 : 0;
  }
  return g_eager == 1;
}

Stream* S(cudaStream_t s) { return s ? reinterpret_cast<Stream*>(s) : &g_default; }
Event* E(cudaEvent_t e) { return reinterpret_cast<Event*>(e); }

std::vector<Stream*> all_streams() {
  std::vector<Stream*> v{&g_default};
  for (Stream* s : g_streams)
    if (s->alive) v.push_back(s);
  return v;
}

// run the head operation of `s` if it can complete
bool try_head(Stream* s) {
  if (s->busy || s->q.empty()) return false;
  s->busy = true;
  const bool done = s->q.front().run();
  s->busy = false;
  if (done) {
    s->q.pop_front();
    ++g_ops_run;
  }
  return done;
}

bool advance_others_impl(Stream* self) {
  for (Stream* s : all_streams())
    if (s != self && try_head(s)) return true;
  return false;
}

// run `s` until its queue is empty
cudaError_t drain(Stream* s) {
  while (!s->q.empty()) {
    if (s->busy) {  // a kernel of this very stream is synchronising on it: cannot happen on a real device either
      fakecuda::error("cudaStreamSynchronize on a stream from inside one of its own operations");
      return cudaErrorUnknown;
    }
    if (try_head(s)) continue;
    if (!advance_others_impl(s)) {
      fakecuda::error("deadlock: the head operation of a synchronised stream is blocked and no other stream can run");
      s->q.clear();
      return cudaErrorLaunchFailure;
    }
  }
  return cudaSuccess;
}

void pump_all() {
  for (bool progress = true; progress;) {
    progress = false;
    for (Stream* s : all_streams())
      while (try_head(s)) progress = true;
  }
}

void* block_alloc(size_t size) {
  void* p = malloc(size ? size : 1);
  if (!p) return nullptr;
  memset(p, 0xCD, size);
  g_blocks[reinterpret_cast<uintptr_t>(p)] = size;
  return p;
}

bool block_free(void* p) {
  auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
  if (it == g_blocks.end()) return false;
  g_blocks.erase(it);
  free(p);
  return true;
}

// the block that contains p, or end()
std::map<uintptr_t, size_t>::const_iterator find_block(const void* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  auto it = g_blocks.upper_bound(a);
  if (it == g_blocks.begin()) return g_blocks.end();
  --it;
  return a < it->first + it->second ? it : g_blocks.end();
}

}  // namespace

namespace fakecuda {

void error(const std::string& what) {
  if (g_errors.size() < 4000) g_errors += what + "\n";
}

bool in_bounds(const void* p, size_t n) {
  if (n == 0) return true;
  auto it = find_block(p);
  if (it == g_blocks.end()) return false;
  return reinterpret_cast<uintptr_t>(p) + n <= it->first + it->second;
}

bool check(const void* p, size_t n, const char* what) {
  if (in_bounds(p, n)) return true;
  char buf[256];
  snprintf(buf, sizeof buf, "%s: [%p, +%zu) is not inside one device / pinned block", what, p, n);
  error(buf);
  return false;
}

cudaError_t enqueue(cudaStream_t s, OpFn fn, bool reads_host) {
  Stream* st = S(s);
  st->q.push_back(Op{std::move(fn), reads_host});
  if (eager()) pump_all();
  return cudaSuccess;
}

bool advance_others(cudaStream_t self) { return advance_others_impl(S(self)); }

}  // namespace fakecuda

// a range that starts inside a tracked block must end inside it; untracked memory is the caller's (NumPy arrays)
static bool range_ok(const void* p, size_t n, const char* what) {
  if (n == 0 || find_block(p) == g_blocks.end()) return true;
  return fakecuda::check(p, n, what);
}

extern "C" {

// ---- introspection for the tests ----
const char* fake_cuda_errors(void) { return g_errors.c_str(); }
void fake_cuda_clear_errors(void) { g_errors.clear(); }
unsigned long long fake_cuda_ops_run(void) { return g_ops_run; }
unsigned long long fake_cuda_h2d_bytes(void) { return g_bytes_h2d; }
void fake_cuda_reset_counters(void) { g_ops_run = 0, g_bytes_h2d = 0; }
// host->device copies still queued on any stream: must be 0 whenever an ABI call has returned
int fake_cuda_pending_host_reads(void) {
  int n = 0;
  for (Stream* s : all_streams())
    for (const Op& o : s->q) n += o.reads_host ? 1 : 0;
  return n;
}
int fake_cuda_pending_ops(void) {
  int n = 0;
  for (Stream* s : all_streams()) n += (int)s->q.size();
  return n;
}
int fake_cuda_live_blocks(void) { return (int)g_blocks.size(); }
// forget everything that is still queued (after a failed scenario the queued copies may point at arrays that are gone)
void fake_cuda_drop_pending(void) {
  for (Stream* s : all_streams()) s->q.clear();
}

// ---- the runtime entry points plsvo_abi.cu links against ----
cudaError_t cudaGetDeviceCount(int* count) {
  *count = 1;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int device) { return device == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr attr, int) {
  if (attr == cudaDevAttrMultiProcessorCount) *value = 148;
  else if (attr == cudaDevAttrMaxSharedMemoryPerBlockOptin) *value = 232448;  // 227 KB, sm_100
  else return cudaErrorInvalidValue;
  return cudaSuccess;
}
const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaErrorModel"; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "error reported by the model runtime"; }

cudaError_t cudaMalloc(void** p, size_t size) {
  *p = block_alloc(size);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
// cudaFree and cudaFreeHost synchronise the device before they release the block (everything queued that can run, runs)
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  pump_all();
  if (!block_free(p)) {
    fakecuda::error("cudaFree of a pointer that is not a live device block");
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t size, unsigned int) {
  *p = block_alloc(size);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) {
  if (!p) return cudaSuccess;
  pump_all();
  if (!block_free(p)) {
    fakecuda::error("cudaFreeHost of a pointer that is not a live pinned block");
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) {
  Stream* st = new Stream();
  g_streams.push_back(st);
  *s = reinterpret_cast<cudaStream_t>(st);
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  Stream* st = S(s);
  drain(st);  // CUDA lets queued work finish
  st->alive = false;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { return drain(S(s)); }

cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) {
  *e = reinterpret_cast<cudaEvent_t>(new Event());
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  // operations that wait on the event keep a raw pointer: leak the few bytes instead of tracking them
  (void)e;
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  Event* ev = E(e);
  const uint64_t gen = ++ev->gen_enqueued;
  return fakecuda::enqueue(s, [ev, gen]() {
    if (ev->gen_done < gen) ev->gen_done = gen;
    return true;
  });
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int) {
  Event* ev = E(e);
  const uint64_t gen = ev->gen_enqueued;  // the most recent record at the time of this call (0: never recorded)
  return fakecuda::enqueue(s, [ev, gen]() { return ev->gen_done >= gen; });
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  Event* ev = E(e);
  const uint64_t gen = ev->gen_enqueued;
  while (ev->gen_done < gen)
    if (!advance_others_impl(nullptr)) {
      fakecuda::error("deadlock in cudaEventSynchronize: the record can never run");
      return cudaErrorLaunchFailure;
    }
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  if (E(a)->gen_done < E(a)->gen_enqueued || E(b)->gen_done < E(b)->gen_enqueued) return cudaErrorNotReady;
  *ms = 0.001f;
  return cudaSuccess;
}

cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, enum cudaMemcpyKind kind, cudaStream_t s) {
  if (!range_ok(dst, n, "cudaMemcpyAsync destination") || !range_ok(src, n, "cudaMemcpyAsync source")) return cudaErrorInvalidValue;
  if (kind != cudaMemcpyDeviceToHost && !fakecuda::in_bounds(dst, n) && n) {
    fakecuda::error("cudaMemcpyAsync: destination of a host->device / device->device copy is not device memory");
    return cudaErrorInvalidValue;
  }
  const bool h2d = kind == cudaMemcpyHostToDevice;
  const bool dst_tracked = fakecuda::in_bounds(dst, n), src_tracked = fakecuda::in_bounds(src, n);
  return fakecuda::enqueue(s, [=]() {
    // cudaFree waits for everything that can run; a copy that was still blocked then has lost its block
    if ((dst_tracked && !fakecuda::in_bounds(dst, n)) || (src_tracked && !fakecuda::in_bounds(src, n))) {
      fakecuda::error("a queued copy ran after one of its device / pinned blocks had been freed");
      return true;
    }
    memmove(dst, src, n);
    if (h2d) g_bytes_h2d += n;
    return true;
  }, h2d);
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t s) {
  if (width > dpitch || width > spitch) return cudaErrorInvalidPitchValue;
  const size_t dspan = height ? (height - 1) * dpitch + width : 0, sspan = height ? (height - 1) * spitch + width : 0;
  if (!range_ok(dst, dspan, "cudaMemcpy2DAsync destination") || !range_ok(src, sspan, "cudaMemcpy2DAsync source"))
    return cudaErrorInvalidValue;
  const bool h2d = kind == cudaMemcpyHostToDevice;
  const bool dst_tracked = fakecuda::in_bounds(dst, dspan), src_tracked = fakecuda::in_bounds(src, sspan);
  return fakecuda::enqueue(s, [=]() {
    if ((dst_tracked && !fakecuda::in_bounds(dst, dspan)) || (src_tracked && !fakecuda::in_bounds(src, sspan))) {
      fakecuda::error("a queued 2-D copy ran after one of its device / pinned blocks had been freed");
      return true;
    }
    for (size_t y = 0; y < height; ++y) memmove(static_cast<char*>(dst) + y * dpitch, static_cast<const char*>(src) + y * spitch, width);
    if (h2d) g_bytes_h2d += width * height;
    return true;
  }, h2d);
}
cudaError_t cudaMemsetAsync(void* p, int value, size_t n, cudaStream_t s) {
  if (!fakecuda::check(p, n, "cudaMemsetAsync")) return cudaErrorInvalidValue;
  return fakecuda::enqueue(s, [p, value, n]() {
    if (!fakecuda::check(p, n, "queued cudaMemsetAsync (block freed meanwhile?)")) return true;
    memset(p, value, n);
    return true;
  });
}

}  // extern "C"
