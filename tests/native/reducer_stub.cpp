// Test double of the CUDA runtime + collective entry points used by ops/csrc/tfy_reducer.cpp, so that the native
// DDP reducer (gradient-ready counting, strictly ordered bucket launches, event / stream choreography) can be
// exercised on a CPU-only box: every call is appended to a log the test reads back.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include <string>

#include "tfy_common.cuh"

static std::string g_log;
static int g_fail_bucket_off = -1;     // tfy_allreduce fails for the bucket at this offset (error propagation test)
static uintptr_t g_next_handle = 0x1000;

static void logf(const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_log += buf;
    g_log += "\n";
}

extern "C" {

cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { *e = (cudaEvent_t)(g_next_handle += 16); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { logf("record ev=%p stream=%p", (void*)e, (void*)s); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int) { logf("wait stream=%p ev=%p", (void*)s, (void*)e); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned int, int prio) { *s = (cudaStream_t)0xC0; logf("comm stream prio=%d", prio); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }

int tfy_allreduce(const TfyCommCtx*, int dtype, int algo, uint64_t off, size_t n, float scale, void*, int, int, cudaStream_t s) {
    logf("allreduce off=%llu n=%zu dtype=%d algo=%d scale=%.4f stream=%p", (unsigned long long)off, n, dtype, algo, scale, (void*)s);
    return (int)off == g_fail_bucket_off ? -7 : 0;
}

int tfy_fused_step_shard_range(const TfyCommCtx*, int, int, int opt, int, uint64_t grad_off, uint64_t param_off, size_t shard_n,
                               float*, float*, float*, TfyOptHyper*, int zero_grads, int, int, size_t, size_t, int advance,
                               cudaStream_t s) {
    logf("fused off=%llu poff=%llu shard=%zu opt=%d zero=%d advance=%d stream=%p", (unsigned long long)grad_off,
         (unsigned long long)param_off, shard_n, opt, zero_grads, advance, (void*)s);
    return 0;
}

int stub_log(char* out, int cap) {
    int n = (int)g_log.size() < cap - 1 ? (int)g_log.size() : cap - 1;
    memcpy(out, g_log.data(), n);
    out[n] = 0;
    return n;
}
void stub_reset(int fail_off) { g_log.clear(); g_fail_bucket_off = fail_off; }

}  // extern "C"
