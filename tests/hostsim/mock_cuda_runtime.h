// mock_cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  A small in-process imitation of the part of the CUDA runtime that
// fourq_b200/csrc/capi.cu uses, so that the host engine (slices, chunk schedule, staging, the per-GPU feeder / drainer
// threads, error paths) can be exercised -- also under ThreadSanitizer / AddressSanitizer -- on a machine without a GPU.
// Streams are worker threads that execute their queue in order; "device memory" is host memory tracked in a registry;
// "kernels" are the CPU instruction-level simulation of the device code (mock_kernels.cpp -> hostsim.cpp).
// Only tests/hostsim builds with -DFQ_MOCK_CUDA include this; the product always builds against the real <cuda_runtime.h>.
#pragma once
#include <cstddef>
#include <functional>

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
struct cudaDeviceProp { int multiProcessorCount; };
struct MockStream; struct MockEvent;
typedef MockStream* cudaStream_t;
typedef MockEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaHostAllocPortable = 1, cudaEventDisableTiming = 2, cudaHostRegisterPortable = 1 };

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int dev);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void* p);

// for the mock kernels: run `fn` on the stream's thread after everything enqueued before it
void mock_stream_enqueue(cudaStream_t s, std::function<void()> fn);
