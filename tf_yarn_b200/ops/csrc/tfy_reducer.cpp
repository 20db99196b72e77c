// Native bucket reducer of the DDP-compatible wrapper (tf_yarn_b200/parallel/ddp.py).
//
// torch's DistributedDataParallel keeps its per-bucket bookkeeping in C++ (c10d::Reducer): gradient-ready
// counting, bucket launch in a FIXED order, a communication stream with events on both sides.  This is the
// same piece for the hand-written collectives: the Python autograd hook makes ONE call per parameter
// (tfy_reducer_mark_ready); counting, ordering, event record / stream wait and the kernel launch happen here.
//   * plain mode : bucket i is all-reduced in place (NVLS multimem.ld_reduce + multimem.st, or two-shot P2P),
//                  averaged, on the reducer's communication stream while backward continues;
//   * fused mode : bucket i runs the fused reduce-scatter -> optimizer -> all-gather kernel (K4) instead, so
//                  the optimizer step also overlaps backward and `optimizer.step()` has nothing left to do
//                  (DistributedDataParallel.fuse_optimizer).
// Buckets are launched strictly in index order (bucket i only after 0..i-1): the kernels of different ranks
// are paired by launch order, so every rank must issue the same sequence even when autograd completes the
// buckets in a different order on some rank.  (reference path being replaced: torch DDP over NCCL as the
// reference's worker uses it, tf_yarn/pytorch/tasks/worker.py:105-107, experiment.py:23-27.)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "tfy_common.cuh"

extern "C" {
int tfy_allreduce(const TfyCommCtx* c, int dtype, int algo, uint64_t off, size_t n, float scale, void* out, int grid,
                  int block, cudaStream_t s);
int tfy_fused_step_shard_range(const TfyCommCtx* c, int grad_dtype, int param_dtype, int opt, int mode,
                               uint64_t grad_off, uint64_t param_off, size_t shard_n, float* master, float* s1,
                               float* s2, TfyOptHyper* hp, int zero_grads, int grid, int block, size_t g0, size_t g1,
                               int advance, cudaStream_t s);
}

namespace {

struct Bucket {
    uint64_t grad_off = 0;
    size_t n = 0;            // elements (padded so that it splits into world x 16-byte packs)
    int dtype = 0;
    int n_params = 0;
    int pending = 0;
    cudaEvent_t ready = nullptr, done = nullptr;
    // fused mode
    bool fused = false;
    uint64_t param_off = 0;
    float *master = nullptr, *s1 = nullptr, *s2 = nullptr;
    size_t shard_n = 0;
    int param_dtype = 0, opt = 0, mode = 0;
    TfyOptHyper* hp = nullptr;
};

struct Reducer {
    TfyCommCtx ctx;
    std::vector<Bucket> buckets;
    std::vector<int> param_bucket;
    int next = 0;              // buckets [0, next) have been launched in this backward
    int last_launched = -1;
    cudaStream_t comm = nullptr;
    int algo = 2;
    long launches = 0;
    int last_error = 0;

    int launch(int b, cudaStream_t cur) {
        Bucket& k = buckets[b];
        cudaEventRecord(k.ready, cur);
        cudaStreamWaitEvent(comm, k.ready, 0);
        int rc;
        bool launched = true;
        if (k.fused) {
            // zero_grads = 1: the kernel clears the bucket behind itself (autograd accumulates into it next step)
            rc = tfy_fused_step_shard_range(&ctx, k.dtype, k.param_dtype, k.opt, k.mode, k.grad_off, k.param_off,
                                            k.shard_n, k.master, k.s1, k.s2, k.hp, 1, 0, 0, 0, (size_t)-1, 1, comm);
        } else if (ctx.world > 1) {
            rc = tfy_allreduce(&ctx, k.dtype, algo, k.grad_off, k.n, 1.0f / (float)ctx.world, nullptr, 0, 0, comm);
        } else {
            rc = 0;                  // one rank, no optimizer fused: nothing to exchange
            launched = false;
        }
        cudaEventRecord(k.done, comm);
        if (rc == 0 && launched) ++launches;
        else last_error = rc;
        last_launched = b;
        return rc;
    }

    int launch_ready_prefix(cudaStream_t cur) {
        while (next < (int)buckets.size() && buckets[next].pending == 0) {
            const int rc = launch(next, cur);
            ++next;
            if (rc != 0) return rc;
        }
        return 0;
    }
};

}  // namespace

extern "C" {

void tfy_reducer_destroy(void* h);

// offs / ns / dtypes / n_params: one entry per bucket; param_bucket: bucket index of every parameter.
void* tfy_reducer_create(const TfyCommCtx* ctx, int n_buckets, const uint64_t* offs, const size_t* ns, const int* dtypes,
                         const int* n_params, int n_total_params, const int* param_bucket, int algo) {
    auto* r = new Reducer();
    r->ctx = *ctx;
    r->algo = algo;
    r->buckets.resize(n_buckets);
    for (int b = 0; b < n_buckets; ++b) {
        Bucket& k = r->buckets[b];
        k.grad_off = offs[b];
        k.n = ns[b];
        k.dtype = dtypes[b];
        k.n_params = n_params[b];
        k.pending = n_params[b];
        if (cudaEventCreateWithFlags(&k.ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming) != cudaSuccess) {
            tfy_reducer_destroy(r);      // also releases the events created so far
            return nullptr;
        }
    }
    r->param_bucket.assign(param_bucket, param_bucket + n_total_params);
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&r->comm, cudaStreamNonBlocking, hi) != cudaSuccess) {
        r->comm = nullptr;
        tfy_reducer_destroy(r);
        return nullptr;
    }
    return r;
}

int tfy_reducer_set_fused(void* h, int bucket, int param_dtype, uint64_t param_off, float* master, float* s1, float* s2,
                          size_t shard_n, int opt, int mode, TfyOptHyper* hp) {
    auto* r = (Reducer*)h;
    if (!r || bucket < 0 || bucket >= (int)r->buckets.size()) return -2;
    Bucket& k = r->buckets[bucket];
    k.fused = true;
    k.param_dtype = param_dtype;
    k.param_off = param_off;
    k.master = master; k.s1 = s1; k.s2 = s2;
    k.shard_n = shard_n;
    k.opt = opt; k.mode = mode; k.hp = hp;
    return 0;
}

// called from the post-accumulate-grad hook of parameter `param`; `cur` = the stream backward runs on
int tfy_reducer_mark_ready(void* h, int param, cudaStream_t cur) {
    auto* r = (Reducer*)h;
    if (!r || param < 0 || param >= (int)r->param_bucket.size()) return -2;
    Bucket& k = r->buckets[r->param_bucket[param]];
    if (k.pending > 0 && --k.pending == 0) return r->launch_ready_prefix(cur);
    return 0;
}

// end of backward: launch what is left (in order; parameters that got no gradient contribute zeros), make the
// compute stream wait for the last collective, re-arm the counters
int tfy_reducer_finalize(void* h, cudaStream_t cur) {
    auto* r = (Reducer*)h;
    int rc = 0;
    for (; r->next < (int)r->buckets.size(); ++r->next) {
        const int e = r->launch(r->next, cur);
        if (e != 0 && rc == 0) rc = e;
    }
    if (r->last_launched >= 0) cudaStreamWaitEvent(cur, r->buckets[r->last_launched].done, 0);
    for (auto& k : r->buckets) k.pending = k.n_params;
    r->next = 0;
    r->last_launched = -1;
    return rc;
}

long tfy_reducer_launches(void* h) { return h ? ((Reducer*)h)->launches : 0; }
void* tfy_reducer_stream(void* h) { return h ? (void*)((Reducer*)h)->comm : nullptr; }
int tfy_reducer_next(void* h) { return h ? ((Reducer*)h)->next : -1; }

void tfy_reducer_destroy(void* h) {
    auto* r = (Reducer*)h;
    if (!r) return;
    for (auto& k : r->buckets) {
        if (k.ready) cudaEventDestroy(k.ready);
        if (k.done) cudaEventDestroy(k.done);
    }
    if (r->comm) cudaStreamDestroy(r->comm);
    delete r;
}

}  // extern "C"
