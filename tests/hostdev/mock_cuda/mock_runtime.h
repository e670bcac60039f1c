// TEST INFRASTRUCTURE - not part of the product, never linked into it, never loaded unless a test names it.
//
// A synchronous stand-in for the CUDA runtime API that csrc/ecmgpu.cu (the host side of the C ABI) uses, so that the
// REAL host glue - buffer sizing, launch order, mode switches, the strip phases, the query and planning entry points -
// can be driven through the real C ABI on a machine without a GPU: device memory is host memory, every "asynchronous"
// call completes before it returns, streams and events are tokens, kernels run at once under the SIMT emulator
// (simt.h) with the grid and block sizes the glue asked for.  tests/hostdev/mock_cuda/make_mock.py rewrites the
// `kernel<<<grid, block, shmem, stream>>>(args)` launches of ecmgpu.cu into hd_mock_launch(grid, block, kernel, args).
// CUDA graphs: a capture records the enqueued operations as closures, a graph launch replays them.
// What this CANNOT show: anything about timing or overlap, the peer / IPC transports, or the real library sort.  The product library itself has no CPU path: without a device it refuses to run.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct HdStream* cudaStream_t;
struct HdEvent { double t_ms; };  // an event is the wall-clock time of its record: elapsed times are host times
typedef struct HdEvent* cudaEvent_t;
// A captured graph is the list of operations enqueued between BeginCapture and EndCapture, as closures that own copies
// of their arguments (like a real graph bakes kernel parameters at capture); a launch runs them in order.
struct HdGraph { std::vector<std::function<void()>> ops; };
typedef HdGraph* cudaGraph_t;
typedef HdGraph* cudaGraphExec_t;
inline thread_local HdGraph* hd_capturing = nullptr;  // per OS thread (one rank per thread in the multi-rank tests)
template <class F>
static inline void hd_enqueue(F&& op) {
    if (hd_capturing) hd_capturing->ops.push_back(std::function<void()>(op));
    else op();
}
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaStreamCaptureModeRelaxed = 2, cudaIpcMemLazyEnablePeerAccess = 1 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "not supported by the mock runtime" : "mock runtime error"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = (size_t)2 << 30; *total_b = (size_t)4 << 30; return cudaSuccess; }
// fresh device memory is POISONED: cudaMalloc promises nothing about its content, and a kernel that reads what nobody
// wrote should fail here rather than pass because a fresh GPU allocation happened to be zero
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, n);
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { hd_enqueue([=] { memmove(d, s, n); }); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { hd_enqueue([=] { memmove(d, s, n); }); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { hd_enqueue([=] { memset(d, v, n); }); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline double hd_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new HdEvent{hd_now_ms()}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new HdEvent{hd_now_ms()}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
// captured like every other stream operation: an event-record node stamps the event at every graph launch
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { hd_enqueue([=] { e->t_ms = hd_now_ms(); }); return cudaSuccess; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
static const unsigned cudaEventRecordExternal = 1u;
static inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = hd_capturing ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone; return cudaSuccess; }
static inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t st, unsigned) { return cudaEventRecord(e, st); }
// every "device" pointer of the mock is host memory the kernels can address
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) { a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = (void*)p; a->hostPointer = (void*)p; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { hd_capturing = new HdGraph; return cudaSuccess; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = hd_capturing; hd_capturing = nullptr; return *g ? cudaSuccess : cudaErrorNotSupported; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new HdGraph(*g); return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& op : e->ops) op(); return cudaSuccess; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
// all "devices" share this process's address space: a handle is the pointer itself
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// kernel<<<grid, block, shmem, stream>>>(args) after make_mock.py
template <class... P, class... A>
static inline void hd_mock_launch(int grid, int block, void (*kernel)(P...), A&&... args) {
    if (grid <= 0 || block <= 0) return;
    // like a real launch: the arguments are evaluated and copied NOW, whenever the kernel runs
    auto bound = std::make_tuple(static_cast<P>(args)...);
    const std::function<void()> fn = [kernel, bound] { std::apply(kernel, bound); };
    hd_enqueue([=] { hd_simt_launch(grid, block, fn); });
}
