// fake_cuda.h — interface between the model CUDA runtime (fake_cudart.cpp) and the model kernels (fake_kernels.cpp).
//
// TEST INFRASTRUCTURE ONLY.  tests/hostmodel/ builds libplsvo_hostmodel.so = the product's own host code
// (pl-svo_b200/csrc/plsvo_abi.cu, compiled unchanged as C++) linked against a single-threaded model of the CUDA runtime
// and against "kernels" that only digest the bytes the real kernels would read.  It exists so that the host pipeline —
// buffer sizing, upload planning, chunking, the arrival gate, stream and event ordering, the frame-chain layout — can be
// exercised in the CPU test tier (tests/test_host_pipeline_cpu.py).  Nothing in the product loads it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <string>

namespace fakecuda {

// One queued stream operation.  run() returns true when the operation has completed and false when it is blocked (an
// event that has not been recorded yet, a gated kernel whose next chunk has not arrived); a blocked operation is retried
// after some operation of another stream has run.
using OpFn = std::function<bool()>;

// Queue `fn` on `s`.  `reads_host` marks host->device copies: none may be pending when an ABI call returns.
cudaError_t enqueue(cudaStream_t s, OpFn fn, bool reads_host = false);

// Run one runnable operation of any stream but `self`.  Returns false when nothing else can make progress.
bool advance_others(cudaStream_t self);

// Record a model error (out-of-bounds access, deadlock, ...); the tests assert that none was recorded.
void error(const std::string& what);

// true when [p, p+n) lies inside one live device / pinned allocation (or n == 0)
bool in_bounds(const void* p, size_t n);
// checks in_bounds and records an error naming `what` otherwise
bool check(const void* p, size_t n, const char* what);

}  // namespace fakecuda
