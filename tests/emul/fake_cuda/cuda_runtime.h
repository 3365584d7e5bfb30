// TEST INFRASTRUCTURE: a stand-in for the few CUDA runtime calls csrc/host_stager.cc makes, so that
// its threading (helper wake-up, piece claiming, ring reuse) can be exercised and race-checked on
// a machine without a GPU (tests/test_host_logic.py builds it with -fsanitize=thread).
// "Device" memory is host memory, copies are synchronous, events are always complete.
#pragma once
#include <stdlib.h>
#include <string.h>
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaMemcpyHostToDevice = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = reinterpret_cast<void*>(1); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
