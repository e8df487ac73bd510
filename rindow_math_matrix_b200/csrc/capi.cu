// capi.cu -- library / device / buffer entry points of the C ABI (include/rindow_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <unordered_set>

#include "common.cuh"

namespace rb200 {

static Context g_ctx;
// allocations made with plain cudaMalloc (CUDA-IPC exportable); everything else comes from the stream-ordered pool
static std::unordered_set<void*> g_shareable;
static thread_local char g_err[1024] = "";

Context& ctx() { return g_ctx; }

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? RB200_E_NOMEM : RB200_E_CUDA;
}

int ensure_workspace(size_t bytes)
{
    Context& c = ctx();
    if (bytes <= c.workspace_bytes) return RB200_OK;
    if (c.workspace) {
        // kernels of a previously adopted stream (rb200_set_stream) may still use the old block
        RB_CUDA(cudaDeviceSynchronize());
        RB_CUDA(cudaFree(c.workspace));
        c.workspace = nullptr;
        c.workspace_bytes = 0;
    }
    size_t want = bytes < (size_t)(8u << 20) ? (size_t)(8u << 20) : bytes;
    RB_CUDA(cudaMalloc(&c.workspace, want));
    c.workspace_bytes = want;
    return RB200_OK;
}

int ensure_counters(int64_t n, int** out)
{
    Context& c = ctx();
    if ((size_t)n > c.counters_len) {
        if (c.counters) {
            RB_CUDA(cudaDeviceSynchronize());
            RB_CUDA(cudaFree(c.counters));
            c.counters = nullptr;
            c.counters_len = 0;
        }
        size_t want = (size_t)n < 16384 ? 16384 : (size_t)n;
        RB_CUDA(cudaMalloc(&c.counters, want * sizeof(int)));
        RB_CUDA(cudaMemsetAsync(c.counters, 0, want * sizeof(int), c.stream));
        c.counters_len = want;
    }
    *out = c.counters;
    return RB200_OK;
}

}  // namespace rb200

using rb200::ctx;

extern "C" {

int rb200_abi_version(void) { return RB200_ABI_VERSION; }

const char* rb200_last_error(void) { return rb200::g_err; }

int rb200_device_count(int* count)
{
    if (!count) RB_INVALID("count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return rb200::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *count = n;
    return RB200_OK;
}

int rb200_init(int device)
{
    rb200::Context& c = ctx();
    if (c.initialised && c.device == device) return RB200_OK;
    if (c.initialised) {
        int rc = rb200_shutdown();
        if (rc != RB200_OK) return rc;
    }
    RB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        rb200::set_error("device %d is sm_%d%d; librindow_b200 contains sm_100a code only (no fallback)",
                         device, prop.major, prop.minor);
        return RB200_E_UNSUPPORTED;
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.cc_major = prop.major;
    c.cc_minor = prop.minor;
    c.total_mem = prop.totalGlobalMem;
    RB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    RB_CUDA(cudaHostAlloc(&c.pinned_scalar, 64, cudaHostAllocDefault));
    {
        // buffers come from the device's stream-ordered memory pool and go back to it without a synchronisation:
        // LinearAlgebra allocates an output per call (src/LinearAlgebra.php:255-260), so cudaMalloc / cudaFree (and the
        // stream synchronise a safe cudaFree needs) would cost more than the kernels.  Keep freed memory cached.
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
        }
    }
    c.launches = 0;
    c.initialised = true;
    return RB200_OK;
}

int rb200_shutdown(void)
{
    rb200::Context& c = ctx();
    if (!c.initialised) return RB200_OK;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    if (c.workspace) cudaFree(c.workspace);
    c.workspace = nullptr;
    c.workspace_bytes = 0;
    if (c.counters) cudaFree(c.counters);
    c.counters = nullptr;
    c.counters_len = 0;
    if (c.gemm_sync) cudaFree(c.gemm_sync);
    c.gemm_sync = nullptr;
    if (c.pinned_scalar) cudaFreeHost(c.pinned_scalar);
    c.pinned_scalar = nullptr;
    if (c.switch_event) cudaEventDestroy(c.switch_event);
    c.switch_event = nullptr;
    if (c.own_stream) cudaStreamDestroy(c.own_stream);
    c.own_stream = nullptr;
    c.stream = nullptr;
    c.initialised = false;
    c.device = -1;
    return RB200_OK;
}

int rb200_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem_bytes)
{
    RB_REQUIRE_INIT();
    rb200::Context& c = ctx();
    if (device) *device = c.device;
    if (sm_count) *sm_count = c.sm_count;
    if (cc_major) *cc_major = c.cc_major;
    if (cc_minor) *cc_minor = c.cc_minor;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)c.total_mem;
    return RB200_OK;
}

int rb200_set_stream(void* cuda_stream)
{
    RB_REQUIRE_INIT();
    rb200::Context& c = ctx();
    // cudaStream_t 0 (legacy default stream) is a valid caller stream; NULL restores ours
    // only when the caller passes the sentinel (void*)-1.
    cudaStream_t next = (cuda_stream == (void*)(intptr_t)-1) ? c.own_stream : (cudaStream_t)cuda_stream;
    if (next != c.stream) {
        // the workspace, the arrival counters and the pinned scalar slot are shared by every launch: work queued on
        // the new stream is ordered after everything queued on the old one
        if (!c.switch_event) RB_CUDA(cudaEventCreateWithFlags(&c.switch_event, cudaEventDisableTiming));
        RB_CUDA(cudaEventRecord(c.switch_event, c.stream));
        RB_CUDA(cudaStreamWaitEvent(next, c.switch_event, 0));
        c.stream = next;
    }
    return RB200_OK;
}

int rb200_get_stream(void** cuda_stream)
{
    RB_REQUIRE_INIT();
    if (!cuda_stream) RB_INVALID("cuda_stream is NULL");
    *cuda_stream = (void*)ctx().stream;
    return RB200_OK;
}

int rb200_sync(void)
{
    RB_REQUIRE_INIT();
    RB_CUDA(cudaStreamSynchronize(ctx().stream));
    return RB200_OK;
}

int rb200_set_default_gemm_mode(int mode)
{
    if (mode < RB200_GEMM_FP32_SIMT || mode > RB200_GEMM_TF32X3) RB_INVALID("unknown gemm mode %d", mode);
    ctx().default_gemm_mode = mode;
    return RB200_OK;
}

int rb200_get_default_gemm_mode(int* mode)
{
    if (!mode) RB_INVALID("mode is NULL");
    *mode = ctx().default_gemm_mode;
    return RB200_OK;
}

int rb200_launch_count(int64_t* count, int reset)
{
    if (count) *count = ctx().launches;
    if (reset) ctx().launches = 0;
    return RB200_OK;
}

const char* rb200_last_gemm_path(void) { return ctx().last_gemm_path; }

/* ---- buffers ------------------------------------------------------------------------ */

int rb200_buffer_alloc(int64_t bytes, void** dptr)
{
    RB_REQUIRE_INIT();
    if (!dptr) RB_INVALID("dptr is NULL");
    if (bytes < 0) RB_INVALID("negative size");
    *dptr = nullptr;
    if (bytes == 0) bytes = 16;
    static const int use_pool = getenv("RB200_POOL") ? atoi(getenv("RB200_POOL")) : 1;
    if (!use_pool) return rb200_buffer_alloc_shareable(bytes, dptr);
    RB_CUDA(cudaMallocAsync(dptr, (size_t)bytes, ctx().stream));
    return RB200_OK;
}

int rb200_buffer_alloc_shareable(int64_t bytes, void** dptr)
{
    RB_REQUIRE_INIT();
    if (!dptr) RB_INVALID("dptr is NULL");
    if (bytes < 0) RB_INVALID("negative size");
    *dptr = nullptr;
    if (bytes == 0) bytes = 16;
    RB_CUDA(cudaMalloc(dptr, (size_t)bytes));
    rb200::g_shareable.insert(*dptr);
    return RB200_OK;
}

int rb200_buffer_free(void* dptr)
{
    RB_REQUIRE_INIT();
    if (!dptr) return RB200_OK;
    auto it = rb200::g_shareable.find(dptr);
    if (it != rb200::g_shareable.end()) {
        rb200::g_shareable.erase(it);
        RB_CUDA(cudaStreamSynchronize(ctx().stream));
        RB_CUDA(cudaFree(dptr));
        return RB200_OK;
    }
    // stream-ordered: kernels already queued on the library stream keep the memory until they have run
    RB_CUDA(cudaFreeAsync(dptr, ctx().stream));
    return RB200_OK;
}

int rb200_buffer_h2d_async(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes)
{
    RB_REQUIRE_INIT();
    if (!dptr || !host || bytes < 0 || dst_offset_bytes < 0) RB_INVALID("bad h2d arguments");
    if (bytes == 0) return RB200_OK;
    RB_CUDA(cudaMemcpyAsync((char*)dptr + dst_offset_bytes, host, (size_t)bytes,
                            cudaMemcpyHostToDevice, ctx().stream));
    return RB200_OK;
}

int rb200_buffer_d2h_async(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes)
{
    RB_REQUIRE_INIT();
    if (!dptr || !host || bytes < 0 || src_offset_bytes < 0) RB_INVALID("bad d2h arguments");
    if (bytes == 0) return RB200_OK;
    RB_CUDA(cudaMemcpyAsync(host, (const char*)dptr + src_offset_bytes, (size_t)bytes,
                            cudaMemcpyDeviceToHost, ctx().stream));
    return RB200_OK;
}

int rb200_buffer_h2d(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes)
{
    int rc = rb200_buffer_h2d_async(dptr, dst_offset_bytes, host, bytes);
    if (rc != RB200_OK) return rc;
    RB_CUDA(cudaStreamSynchronize(ctx().stream));
    return RB200_OK;
}

int rb200_buffer_d2h(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes)
{
    int rc = rb200_buffer_d2h_async(host, dptr, src_offset_bytes, bytes);
    if (rc != RB200_OK) return rc;
    RB_CUDA(cudaStreamSynchronize(ctx().stream));
    return RB200_OK;
}

int rb200_buffer_d2d(void* dst, int64_t dst_offset_bytes, const void* src, int64_t src_offset_bytes,
                     int64_t bytes)
{
    RB_REQUIRE_INIT();
    if (!dst || !src || bytes < 0 || dst_offset_bytes < 0 || src_offset_bytes < 0) RB_INVALID("bad d2d arguments");
    if (bytes == 0) return RB200_OK;
    RB_CUDA(cudaMemcpyAsync((char*)dst + dst_offset_bytes, (const char*)src + src_offset_bytes,
                            (size_t)bytes, cudaMemcpyDeviceToDevice, ctx().stream));
    return RB200_OK;
}

/* ---- peer memory (one process per GPU; NVLink / NVSwitch peer access through CUDA IPC) -------------- */

int rb200_ipc_export(void* dptr, void* handle64)
{
    RB_REQUIRE_INIT();
    if (!dptr || !handle64) RB_INVALID("bad ipc_export arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == RB200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    RB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, sizeof(h));
    return RB200_OK;
}

int rb200_ipc_open(const void* handle64, void** dptr)
{
    RB_REQUIRE_INIT();
    if (!dptr || !handle64) RB_INVALID("bad ipc_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *dptr = nullptr;
    RB_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RB200_OK;
}

int rb200_ipc_close(void* dptr)
{
    RB_REQUIRE_INIT();
    if (!dptr) return RB200_OK;
    RB_CUDA(cudaStreamSynchronize(ctx().stream));
    RB_CUDA(cudaIpcCloseMemHandle(dptr));
    return RB200_OK;
}

int rb200_host_alloc(int64_t bytes, void** hptr)
{
    RB_REQUIRE_INIT();
    if (!hptr || bytes < 0) RB_INVALID("bad host_alloc arguments");
    if (bytes == 0) bytes = 16;
    RB_CUDA(cudaHostAlloc(hptr, (size_t)bytes, cudaHostAllocDefault));
    return RB200_OK;
}

int rb200_host_free(void* hptr)
{
    RB_REQUIRE_INIT();
    if (!hptr) return RB200_OK;
    RB_CUDA(cudaFreeHost(hptr));
    return RB200_OK;
}

/* ---- CUDA graphs ------------------------------------------------------------------------------- */

namespace {
struct GraphExec { cudaGraphExec_t exec; int64_t launches; };
int64_t g_capture_launch_base = -1;
}

int rb200_graph_begin(void)
{
    RB_REQUIRE_INIT();
    if (g_capture_launch_base >= 0) RB_INVALID("a capture is already in progress");
    RB_CUDA(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeRelaxed));
    g_capture_launch_base = ctx().launches;
    return RB200_OK;
}

int rb200_graph_end(void** graph_exec)
{
    RB_REQUIRE_INIT();
    if (!graph_exec) RB_INVALID("graph_exec is NULL");
    if (g_capture_launch_base < 0) RB_INVALID("no capture in progress");
    const int64_t recorded = ctx().launches - g_capture_launch_base;
    ctx().launches = g_capture_launch_base;              // recorded, not executed
    g_capture_launch_base = -1;
    cudaGraph_t graph = nullptr;
    RB_CUDA(cudaStreamEndCapture(ctx().stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return rb200::cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    GraphExec* g = new GraphExec{exec, recorded};
    *graph_exec = g;
    return RB200_OK;
}

int rb200_graph_launch(void* graph_exec)
{
    RB_REQUIRE_INIT();
    if (!graph_exec) RB_INVALID("graph_exec is NULL");
    GraphExec* g = reinterpret_cast<GraphExec*>(graph_exec);
    RB_CUDA(cudaGraphLaunch(g->exec, ctx().stream));
    rb200::count_launch((int)g->launches);
    return RB200_OK;
}

int rb200_graph_destroy(void* graph_exec)
{
    RB_REQUIRE_INIT();
    if (!graph_exec) return RB200_OK;
    GraphExec* g = reinterpret_cast<GraphExec*>(graph_exec);
    RB_CUDA(cudaStreamSynchronize(ctx().stream));
    RB_CUDA(cudaGraphExecDestroy(g->exec));
    delete g;
    return RB200_OK;
}

/* ---- events ------------------------------------------------------------------------------------- */

int rb200_event_create(void** event)
{
    RB_REQUIRE_INIT();
    if (!event) RB_INVALID("event is NULL");
    cudaEvent_t e = nullptr;
    RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *event = (void*)e;
    return RB200_OK;
}

int rb200_event_destroy(void* event)
{
    RB_REQUIRE_INIT();
    if (!event) return RB200_OK;
    RB_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return RB200_OK;
}

int rb200_event_record(void* event)
{
    RB_REQUIRE_INIT();
    if (!event) RB_INVALID("event is NULL");
    RB_CUDA(cudaEventRecord((cudaEvent_t)event, ctx().stream));
    return RB200_OK;
}

int rb200_event_synchronize(void* event)
{
    RB_REQUIRE_INIT();
    if (!event) RB_INVALID("event is NULL");
    RB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return RB200_OK;
}

int rb200_event_query(void* event, int* done)
{
    RB_REQUIRE_INIT();
    if (!event || !done) RB_INVALID("bad event_query arguments");
    cudaError_t e = cudaEventQuery((cudaEvent_t)event);
    if (e == cudaSuccess) { *done = 1; return RB200_OK; }
    if (e == cudaErrorNotReady) { cudaGetLastError(); *done = 0; return RB200_OK; }
    return rb200::cuda_fail(e, "cudaEventQuery", __FILE__, __LINE__);
}

int rb200_event_wait(void* event)
{
    RB_REQUIRE_INIT();
    if (!event) RB_INVALID("event is NULL");
    RB_CUDA(cudaStreamWaitEvent(ctx().stream, (cudaEvent_t)event, 0));
    return RB200_OK;
}

}  // extern "C"
