// Runtime plumbing behind the reference's DLArray / DLStream / DLEvent C API
// (src/common/c_runtime_api.h:28-77; reference implementation src/common/c_runtime_api.cc,
// src/cuda_common/*).  Only what the embedding hot path's callers need.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>

#include <algorithm>

#include "hb_rows.cuh"

namespace hb {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_last_error(const std::string &msg) {
    t_last_error = msg;
}

static std::atomic<u32> g_hot_threshold{0};

u32 default_hot_threshold() {
    u32 t = g_hot_threshold.load();
    if (t == 0) {
        const char *e = getenv("HERALD_HOT_THRESHOLD");
        t = e ? (u32)std::max(1, atoi(e)) : 64u;
        g_hot_threshold.store(t);
    }
    return t;
}

int hot_stages() {
    static const int stages = [] {
        const char *e = getenv("HERALD_HOT_STAGES");
        int v = e ? atoi(e) : kHotStagesDefault;
        return std::min(std::max(v, 2), kHotStagesMax);
    }();
    return stages;
}

u32 ticket_rows() {
    static const u32 rows = [] {
        const char *e = getenv("HERALD_TICKET_ROWS");
        int v = e ? atoi(e) : 8; // the next ticket's work items are prefetched, so tickets can be small
        return (u32)std::min(std::max(v, 4), 32);
    }();
    return rows;
}

u32 medium_threshold() {
    static const u32 rows = [] {
        const char *e = getenv("HERALD_MED_THRESHOLD");
        int v = e ? atoi(e) : (int)kMediumDefault;
        return (u32)std::max(v, 1);
    }();
    return rows;
}

static std::atomic<u64 *> g_seg_trace{nullptr};

u64 *seg_trace_buffer() {
    return g_seg_trace.load(std::memory_order_relaxed);
}

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0)
            sms = 148;
    }
    return sms;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("HERALD_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

__global__ void fill_kernel(float *out, float value, size_t n) {
    pdl_enter();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    // 128-bit stores on the aligned body, scalar tail
    size_t n4 = n / 4;
    float4 v4 = make_float4(value, value, value, value);
    for (size_t k = i; k < n4; k += stride)
        reinterpret_cast<float4 *>(out)[k] = v4;
    for (size_t k = n4 * 4 + i; k < n; k += stride)
        out[k] = value;
}

} // namespace hb

using namespace hb;

extern "C" {

const char *HBGetLastError(void) {
    return t_last_error.c_str();
}

const char *HBVersion(void) {
    return "herald_b200 0.1 sm_100a";
}

uint64_t HBKernelLaunchCount(void) {
    return g_launches.load();
}

int HBSegTraceEnable(int on) {
    HB_API_BEGIN();
    u64 *cur = g_seg_trace.load();
    if (on && !cur) {
        u64 *buf = nullptr;
        HB_CUDA(cudaMalloc((void **)&buf, kTraceWords * sizeof(u64)));
        HB_CUDA(cudaMemset(buf, 0, kTraceWords * sizeof(u64)));
        g_seg_trace.store(buf);
    } else if (!on && cur) {
        g_seg_trace.store(nullptr);
        HB_CUDA(cudaDeviceSynchronize());
        cudaFree(cur);
    }
    HB_API_END();
}

int HBSegTraceRead(unsigned long long *out, size_t words) {
    HB_API_BEGIN();
    u64 *cur = g_seg_trace.load();
    HB_CHECK(cur != nullptr, "segment trace is not enabled");
    HB_CUDA(cudaDeviceSynchronize());
    HB_CUDA(cudaMemcpy(out, cur, std::min(words, kTraceWords) * sizeof(u64), cudaMemcpyDeviceToHost));
    HB_API_END();
}

int HBSetHotThreshold(unsigned rows) {
    g_hot_threshold.store(rows ? rows : 1u);
    return 0;
}

int DLStreamCreate(size_t dev_id, DLStreamHandle *handle) {
    HB_API_BEGIN();
    HB_CUDA(cudaSetDevice((int)dev_id));
    auto *s = new DLStream();
    auto *cs = new cudaStream_t();
    HB_CUDA(cudaStreamCreateWithFlags(cs, cudaStreamNonBlocking));
    s->device_id = (int)dev_id;
    s->handle = cs;
    *handle = s;
    HB_API_END();
}

int DLStreamDestroy(DLStreamHandle handle) {
    HB_API_BEGIN();
    auto *cs = (cudaStream_t *)handle->handle;
    HB_CUDA(cudaStreamDestroy(*cs));
    delete cs;
    delete handle;
    HB_API_END();
}

int DLStreamSync(DLStreamHandle handle) {
    HB_API_BEGIN();
    HB_CUDA(cudaStreamSynchronize(stream_of(handle)));
    HB_API_END();
}

int DLEventCreate(size_t dev_id, DLEventHandle *handle) {
    HB_API_BEGIN();
    HB_CUDA(cudaSetDevice((int)dev_id));
    auto *e = new DLEvent();
    auto *ce = new cudaEvent_t();
    HB_CUDA(cudaEventCreate(ce));
    e->device_id = (int)dev_id;
    e->handle = ce;
    *handle = e;
    HB_API_END();
}

int DLEventDestroy(DLEventHandle handle) {
    HB_API_BEGIN();
    auto *ce = (cudaEvent_t *)handle->handle;
    HB_CUDA(cudaEventDestroy(*ce));
    delete ce;
    delete handle;
    HB_API_END();
}

int DLEventRecord(DLStreamHandle stream_handle, DLEventHandle event_handle) {
    HB_API_BEGIN();
    HB_CUDA(cudaEventRecord(*(cudaEvent_t *)event_handle->handle, stream_of(stream_handle)));
    HB_API_END();
}

int DLEventSync(DLEventHandle handle) {
    HB_API_BEGIN();
    HB_CUDA(cudaEventSynchronize(*(cudaEvent_t *)handle->handle));
    HB_API_END();
}

int DLEventElapsedTime(DLEventHandle start, DLEventHandle ending, float *duration) {
    HB_API_BEGIN();
    HB_CUDA(cudaEventElapsedTime(duration, *(cudaEvent_t *)start->handle,
                                 *(cudaEvent_t *)ending->handle));
    HB_API_END();
}

namespace {
// host arrays allocated without CUDA (no device on the box): freed with free()
std::mutex g_plain_host_mtx;
std::set<void *> g_plain_host;
bool free_plain_host(void *p) {
    std::lock_guard<std::mutex> lock(g_plain_host_mtx);
    auto it = g_plain_host.find(p);
    if (it == g_plain_host.end())
        return false;
    g_plain_host.erase(it);
    free(p);
    return true;
}
} // namespace

int DLArrayAlloc(const index_t *shape, const index_t *stride, index_t ndim, DLContext ctx,
                 DLArrayHandle *out) {
    HB_API_BEGIN();
    HB_CHECK(ndim >= 0 && ndim <= 16, "bad ndim");
    auto *arr = new DLArray();
    arr->ctx = ctx;
    arr->ndim = (int)ndim;
    arr->shape = new int64_t[ndim > 0 ? ndim : 1];
    arr->stride = new int64_t[ndim > 0 ? ndim : 1];
    size_t n = 1;
    for (index_t i = 0; i < ndim; i++) {
        arr->shape[i] = shape[i];
        n *= (size_t)shape[i];
    }
    // row-major strides in elements when the caller gives none
    int64_t run = 1;
    for (index_t i = ndim - 1; i >= 0; i--) {
        arr->stride[i] = stride ? stride[i] : run;
        run *= shape[i];
    }
    size_t bytes = (n ? n : 1) * sizeof(float); // every Hetu tensor is fp32
    arr->data = nullptr;
    cudaError_t err;
    if (ctx.device_type == kGPU) {
        err = cudaSetDevice(ctx.device_id);
        if (err == cudaSuccess)
            err = cudaMalloc(&arr->data, bytes);
    } else {
        // pinned: async H2D/D2H for the cache's host-buffer entry points
        err = cudaHostAlloc(&arr->data, bytes, cudaHostAllocPortable);
        if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) {
            // a box without any CUDA device (host-side planning / data loading only): plain host
            // memory; nothing in this process can launch a kernel or start a copy engine anyway
            (void)cudaGetLastError();
            arr->data = nullptr;
            if (posix_memalign(&arr->data, 64, bytes) == 0) {
                std::lock_guard<std::mutex> lock(g_plain_host_mtx);
                g_plain_host.insert(arr->data);
                err = cudaSuccess;
            }
        }
    }
    if (err != cudaSuccess) {
        delete[] arr->shape;
        delete[] arr->stride;
        delete arr;
        throw Error(std::string("DLArrayAlloc: ") + cudaGetErrorString(err));
    }
    *out = arr;
    HB_API_END();
}

int DLArrayFree(DLArrayHandle handle) {
    HB_API_BEGIN();
    if (handle) {
        if (handle->data) {
            if (handle->ctx.device_type == kGPU)
                HB_CUDA(cudaFree(handle->data));
            else if (!free_plain_host(handle->data))
                HB_CUDA(cudaFreeHost(handle->data));
        }
        delete[] handle->shape;
        delete[] handle->stride;
        delete handle;
    }
    HB_API_END();
}

int DLArrayCopyFromTo(DLArrayHandle from, DLArrayHandle to, DLStreamHandle stream) {
    HB_API_BEGIN();
    size_t n = numel(from);
    HB_CHECK(n == numel(to), "DLArrayCopyFromTo: size mismatch");
    cudaMemcpyKind kind = cudaMemcpyDefault;
    if (from->ctx.device_type != kGPU && to->ctx.device_type != kGPU) {
        std::memcpy(to->data, from->data, n * sizeof(float)); // host to host: no CUDA involved
    } else if (stream) {
        HB_CUDA(cudaMemcpyAsync(to->data, from->data, n * sizeof(float), kind, stream_of(stream)));
    } else {
        HB_CUDA(cudaMemcpy(to->data, from->data, n * sizeof(float), kind));
    }
    HB_API_END();
}

int DLGpuArraySet(DLArrayHandle arr, float value, DLStreamHandle stream_handle) {
    HB_API_BEGIN();
    size_t n = numel(arr);
    if (n) {
        cudaStream_t st = stream_of(stream_handle);
        if (value == 0.f) {
            HB_CUDA(cudaMemsetAsync(arr->data, 0, n * sizeof(float), st));
        } else {
            int blocks = std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)sm_count() * 8);
            HB_LAUNCH(fill_kernel, blocks, 256, 0, st, (float *)arr->data, value, n);
            HB_LAUNCHED();
        }
    }
    HB_API_END();
}

} // extern "C"
