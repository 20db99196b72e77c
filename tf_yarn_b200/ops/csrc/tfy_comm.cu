// Communication kernels over NVLink 5 / NVSwitch peer memory (sm_100a).
//
//   K1  one-shot all-reduce     every rank reads every peer's buffer (latency-optimal, small)
//   K2  two-shot all-reduce     P2P reduce-scatter + P2P all-gather in ONE kernel
//   K3  NVLS all-reduce         multimem.ld_reduce (in-switch sum) + multimem.st (in-switch bcast)
//   K4  fused gradient step     reduce-scatter -> bf16->fp32 cast, 1/N scale -> optimizer update on the
//                               owned fp32 master shard -> all-gather of the updated (bf16|fp32) params,
//                               all inside one kernel; optimizer state is sharded 1/N per rank
//   K7  broadcast               root multicasts (NVLS) or peers pull (P2P)
//   K8  barrier                 see tfy_block_barrier in tfy_common.cuh
//       all-gather              pull-based, used to materialise sharded optimizer state for checkpoints
//
// These replace what the reference delegates to Horovod/gloo, torch-DDP/NCCL
// (reference: tf_yarn/tensorflow/tasks/gloo_allred_task.py:54,72-75,89 and
// tf_yarn/pytorch/tasks/worker.py:101-107).  No NCCL call is made on this path.
#include "tfy_common.cuh"

// ---------------------------------------------------------------------------
// reduction of one 16-byte pack across ranks
// ---------------------------------------------------------------------------
template <typename T, int MODE>
__device__ __forceinline__ void tfy_reduce_pack(const TfyCommCtx& c, uint64_t byte_off, float* f) {
    using P = TfyPack<T>;
    if (MODE == TFY_MODE_LOCAL) {
        P::unpack(tfy_ld16(reinterpret_cast<const void*>(c.peer_base[c.rank] + byte_off)), f);
    } else if (MODE == TFY_MODE_NVLS) {
        P::unpack(P::mc_ld_reduce(reinterpret_cast<const void*>(c.mc_base + byte_off)), f);
    } else {
#pragma unroll
        for (int i = 0; i < P::N; ++i) f[i] = 0.f;
        // fixed rank order => bit-identical result no matter which rank reduces
        for (int r0 = 0; r0 < c.world; r0 += 4) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r0 + j < c.world)
                    v[j] = tfy_ld16_stream(reinterpret_cast<const void*>(c.peer_base[r0 + j] + byte_off));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r0 + j < c.world) P::accum(v[j], f);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void tfy_bcast_pack(const TfyCommCtx& c, uint64_t byte_off, uint4 v) {
    if (MODE == TFY_MODE_LOCAL) {
        tfy_st16(reinterpret_cast<void*>(c.peer_base[c.rank] + byte_off), v);
    } else if (MODE == TFY_MODE_NVLS) {
        tfy_mc_st16(reinterpret_cast<void*>(c.mc_base + byte_off), v);
    } else {
        for (int r = 0; r < c.world; ++r) {
            // start with my own replica +1 so that the N ranks do not all hit the same peer first
            int p = c.rank + 1 + r;
            if (p >= c.world) p -= c.world;
            tfy_st16_sys(reinterpret_cast<void*>(c.peer_base[p] + byte_off), v);
        }
    }
}

// ---------------------------------------------------------------------------
// K1 / K2 / K3   all-reduce
// ---------------------------------------------------------------------------
// `off` is the byte offset of the buffer inside the symmetric arena, identical
// on every rank.  n_packs = number of 16-byte packs (for two-shot / NVLS it must
// be a multiple of world).  Two-shot and NVLS work in place; one-shot writes to
// `out` (a local buffer) because peers may still be reading the input.
template <typename T, int ALGO>
__global__ void __launch_bounds__(512) tfy_allreduce_kernel(TfyCommCtx c, uint64_t off, size_t n_packs, float scale,
                                                            uint4* __restrict__ out) {
    using P = TfyPack<T>;
    tfy_block_barrier(c);  // every rank's input is complete and visible
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (ALGO == TFY_ALGO_ONESHOT) {
        for (size_t i = tid; i < n_packs; i += nthreads) {
            float f[P::N];
            tfy_reduce_pack<T, TFY_MODE_P2P>(c, off + i * 16, f);
#pragma unroll
            for (int k = 0; k < P::N; ++k) f[k] *= scale;
            out[i] = P::pack(f);
        }
    } else {
        const size_t shard = n_packs / c.world;
        const size_t base = shard * c.rank;
        constexpr int MODE = (ALGO == TFY_ALGO_NVLS) ? TFY_MODE_NVLS : TFY_MODE_P2P;
        for (size_t i = tid; i < shard; i += nthreads) {
            float f[P::N];
            const uint64_t bo = off + (base + i) * 16;
            tfy_reduce_pack<T, MODE>(c, bo, f);
#pragma unroll
            for (int k = 0; k < P::N; ++k) f[k] *= scale;
            tfy_bcast_pack<MODE>(c, bo, P::pack(f));
        }
    }
    tfy_block_barrier(c);  // results landed everywhere; inputs may be overwritten
}

// ---------------------------------------------------------------------------
// K7 broadcast / all-gather
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512) tfy_broadcast_kernel(TfyCommCtx c, uint64_t off, size_t n_packs, int root) {
    tfy_block_barrier(c);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (MODE == TFY_MODE_NVLS) {
        if (c.rank == root)
            for (size_t i = tid; i < n_packs; i += nthreads) {
                uint4 v = tfy_ld16(reinterpret_cast<const void*>(c.peer_base[root] + off + i * 16));
                tfy_mc_st16(reinterpret_cast<void*>(c.mc_base + off + i * 16), v);
            }
    } else {
        if (c.rank != root)
            for (size_t i = tid; i < n_packs; i += nthreads) {
                uint4 v = tfy_ld16_stream(reinterpret_cast<const void*>(c.peer_base[root] + off + i * 16));
                tfy_st16(reinterpret_cast<void*>(c.peer_base[c.rank] + off + i * 16), v);
            }
    }
    tfy_block_barrier(c);
}

// buffer layout [world][shard_packs]; rank r owns slice r of its own replica
__global__ void __launch_bounds__(512) tfy_allgather_kernel(TfyCommCtx c, uint64_t off, size_t shard_packs) {
    tfy_block_barrier(c);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (int rr = 1; rr < c.world; ++rr) {
        int r = c.rank + rr;
        if (r >= c.world) r -= c.world;
        const uint64_t bo = off + (uint64_t)r * shard_packs * 16;
        for (size_t i = tid; i < shard_packs; i += nthreads) {
            uint4 v = tfy_ld16_stream(reinterpret_cast<const void*>(c.peer_base[r] + bo + i * 16));
            tfy_st16(reinterpret_cast<void*>(c.peer_base[c.rank] + bo + i * 16), v);
        }
    }
    tfy_block_barrier(c);
}

__global__ void tfy_barrier_kernel(TfyCommCtx c) { tfy_block_barrier(c); }

// ---------------------------------------------------------------------------
// K4  fused reduce-scatter -> optimizer -> all-gather
// ---------------------------------------------------------------------------
template <int OPT>
__device__ __forceinline__ void tfy_opt_update(float& p, float g, float& s1, float& s2, const float lr,
                                               const float p1, const float p2, const float eps, const float wd,
                                               const int flags, const float bc1, const float bc2_rsqrt,
                                               const bool first_step) {
    if (OPT == TFY_OPT_SGD) {
        g += wd * p;
        if (p1 != 0.f) {
            s1 = first_step ? g : p1 * s1 + (1.f - p2) * g;
            g = (flags & 1) ? g + p1 * s1 : s1;
        }
        p -= lr * g;
    } else if (OPT == TFY_OPT_ADADELTA) {
        g += wd * p;
        s1 = p1 * s1 + (1.f - p1) * g * g;                      // E[g^2]
        const float upd = g * sqrtf(s2 + eps) * rsqrtf(s1 + eps);
        s2 = p1 * s2 + (1.f - p1) * upd * upd;                  // E[dx^2]
        p -= lr * upd;
    } else if (OPT == TFY_OPT_ADAM) {
        if (flags & 1) p *= (1.f - lr * wd);                    // AdamW
        else g += wd * p;
        s1 = p1 * s1 + (1.f - p1) * g;
        s2 = p2 * s2 + (1.f - p2) * g * g;
        const float denom = sqrtf(s2) * bc2_rsqrt + eps;
        p -= (lr / bc1) * (s1 / denom);
    } else {  // Adagrad
        g += wd * p;
        s1 += g * g;
        p -= lr * g / (sqrtf(s1) + eps);
    }
}

// GT: gradient dtype in the symmetric grad buffer, PT: dtype of the replicated
// compute parameters.  master/s1/s2: local fp32 arrays of shard_n elements
// (rank r owns elements [r*shard_n, (r+1)*shard_n) of the flat buffers).
template <typename GT, typename PT, int OPT, int MODE>
__global__ void __launch_bounds__(256)
tfy_fused_step_kernel(TfyCommCtx c, uint64_t grad_off, uint64_t param_off, size_t shard_n,
                      float* __restrict__ master, float* __restrict__ s1, float* __restrict__ s2,
                      TfyOptHyper* __restrict__ hp, int zero_grads, size_t e0, size_t e1, int advance) {
    // [e0, e1): element range of the FLAT buffers this launch handles (multiples of 8; the whole buffer for
    // the classic single launch).  Splitting a step into ranges lets the engine run the update of gradients
    // that are final early (the big Dense kernel) on a side stream while backward continues; `advance` is set
    // on the last launch of a step only, so every range of the step sees the same step counter.
    tfy_pdl_sync();
    using GP = TfyPack<GT>;
    using PP = TfyPack<PT>;
    constexpr int NG = 8 / GP::N;  // 16-byte packs per 8 gradient elements
    constexpr int NP = 8 / PP::N;

    const float lr = hp->lr, p1 = hp->p1, p2 = hp->p2, eps = hp->eps, wd = hp->weight_decay;
    const float gscale = hp->grad_scale / (float)c.world;
    const int flags = hp->flags;
    const int step = hp->step;  // completed steps; this launch performs step+1
    float bc1 = 1.f, bc2_rsqrt = 1.f;
    if (OPT == TFY_OPT_ADAM) {
        const float t = (float)(step + 1);
        bc1 = 1.f - powf(p1, t);
        bc2_rsqrt = rsqrtf(1.f - powf(p2, t));
    }
    const bool first_step = (step == 0);

    const size_t shard_start = (MODE == TFY_MODE_LOCAL) ? 0 : shard_n * (size_t)c.rank;
    // part of [e0, e1) that falls into the shard this rank owns, in groups of 8 relative to the shard
    const size_t lo_e = e0 > shard_start ? e0 - shard_start : 0;
    const size_t hi_e = e1 > shard_start ? e1 - shard_start : 0;
    const size_t g_lo = (lo_e < shard_n ? lo_e : shard_n) / 8;
    const size_t groups = (hi_e < shard_n ? hi_e : shard_n) / 8;      // exclusive upper bound
    const size_t tid = g_lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    constexpr bool TWO = (OPT == TFY_OPT_ADADELTA || OPT == TFY_OPT_ADAM);

    // The owned master / optimizer-state shard does not depend on the peers: its loads are issued BEFORE the
    // cross-GPU barrier and before the (2-3 us) in-switch reduction, so their latency hides behind both.
    float4 m0, m1, x0, x1, y0, y1;
    auto load_state = [&](size_t g8) {
        const float4* mp = reinterpret_cast<const float4*>(master + g8 * 8);
        m0 = mp[0]; m1 = mp[1];
        const float4* sp = reinterpret_cast<const float4*>(s1 + g8 * 8);
        x0 = sp[0]; x1 = sp[1];
        if (TWO) {
            const float4* tp = reinterpret_cast<const float4*>(s2 + g8 * 8);
            y0 = tp[0]; y1 = tp[1];
        }
    };
    if (tid < groups) load_state(tid);

    if (MODE != TFY_MODE_LOCAL) tfy_block_barrier(c);  // all ranks finished backward

    for (size_t g8 = tid; g8 < groups; g8 += nthreads) {
        const size_t e = shard_start + g8 * 8;
        if (g8 != tid) load_state(g8);
        float g[8];
#pragma unroll
        for (int k = 0; k < NG; ++k)
            tfy_reduce_pack<GT, MODE>(c, grad_off + e * sizeof(GT) + k * 16, g + k * GP::N);
        float4* mp = reinterpret_cast<float4*>(master + g8 * 8);
        float p[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float b[8];
        if (TWO) {
            b[0] = y0.x; b[1] = y0.y; b[2] = y0.z; b[3] = y0.w;
            b[4] = y1.x; b[5] = y1.y; b[6] = y1.z; b[7] = y1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) b[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            tfy_opt_update<OPT>(p[k], g[k] * gscale, a[k], b[k], lr, p1, p2, eps, wd, flags, bc1, bc2_rsqrt,
                                first_step);
        mp[0] = make_float4(p[0], p[1], p[2], p[3]);
        mp[1] = make_float4(p[4], p[5], p[6], p[7]);
        {
            float4* sp = reinterpret_cast<float4*>(s1 + g8 * 8);
            sp[0] = make_float4(a[0], a[1], a[2], a[3]);
            sp[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
        if (OPT == TFY_OPT_ADADELTA || OPT == TFY_OPT_ADAM) {
            float4* sp = reinterpret_cast<float4*>(s2 + g8 * 8);
            sp[0] = make_float4(b[0], b[1], b[2], b[3]);
            sp[1] = make_float4(b[4], b[5], b[6], b[7]);
        }
#pragma unroll
        for (int k = 0; k < NP; ++k)
            tfy_bcast_pack<MODE>(c, param_off + e * sizeof(PT) + k * 16, PP::pack(p + k * PP::N));
    }

    if (MODE != TFY_MODE_LOCAL) tfy_block_barrier(c);  // updated params visible on every rank

    if (zero_grads) {
        // Clear the gradient replica for the next accumulation.  A CTA may only clear what is known
        // to be consumed: after the exit barrier the same-index CTA of EVERY rank has finished, and
        // those CTAs read (from all replicas, mine included) exactly the groups `tid + k*nthreads`
        // of each rank's shard.  So this thread clears those groups of every shard in the LOCAL
        // replica -- never data another local CTA's peers may still be reading.
        const uint4 z = make_uint4(0, 0, 0, 0);
        char* gb = reinterpret_cast<char*>(c.peer_base[c.rank] + grad_off);
        const int nshards = (MODE == TFY_MODE_LOCAL) ? 1 : c.world;
        for (int r = 0; r < nshards; ++r) {
            for (size_t g8 = tid; g8 < groups; g8 += nthreads) {
                char* p = gb + ((size_t)r * shard_n + g8 * 8) * sizeof(GT);
#pragma unroll
                for (int k = 0; k < NG; ++k) tfy_st16(p + k * 16, z);
            }
        }
    }

    // advance the device-side step counter exactly once per launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&hp->done, 1u);
        if (prev == gridDim.x - 1) {
            hp->done = 0;
            if (advance) hp->step = step + 1;
        }
    }
}

// ---------------------------------------------------------------------------
// host launchers (C ABI, called through ctypes; stream = raw cudaStream_t)
// ---------------------------------------------------------------------------
static inline int tfy_pick_grid(size_t work_items, int block, int max_grid) {
    size_t g = (work_items + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t)max_grid) g = max_grid;
    return (int)g;
}

extern "C" {

int tfy_barrier(const TfyCommCtx* c, int grid, cudaStream_t s) {
    if (grid < 1) grid = 1;
    tfy_barrier_kernel<<<grid, 32, 0, s>>>(*c);
    return (int)cudaGetLastError();
}

// n = number of elements; must be a multiple of (16/sizeof(T)) [* world for algo 1,2]
int tfy_allreduce(const TfyCommCtx* c, int dtype, int algo, uint64_t off, size_t n, float scale, void* out,
                  int grid, int block, cudaStream_t s) {
    const size_t esz = dtype == TFY_BF16 ? 2 : 4;
    const size_t n_packs = n * esz / 16;
    if (n_packs * 16 != n * esz) return -2;
    if (algo != TFY_ALGO_ONESHOT && (n_packs % c->world) != 0) return -3;
    if (algo == TFY_ALGO_ONESHOT && out == nullptr) return -4;
    if (algo == TFY_ALGO_NVLS && c->mc_base == 0) return -5;
    if (block <= 0) block = 512;
    const size_t work = algo == TFY_ALGO_ONESHOT ? n_packs : n_packs / c->world;
    if (grid <= 0) grid = tfy_pick_grid(work, block * 2, 148);
    if (grid > TFY_MAX_BLOCKS) grid = TFY_MAX_BLOCKS;
#define TFY_AR(T, A) tfy_allreduce_kernel<T, A><<<grid, block, 0, s>>>(*c, off, n_packs, scale, (uint4*)out)
    if (dtype == TFY_BF16) {
        if (algo == TFY_ALGO_ONESHOT) TFY_AR(__nv_bfloat16, TFY_ALGO_ONESHOT);
        else if (algo == TFY_ALGO_TWOSHOT) TFY_AR(__nv_bfloat16, TFY_ALGO_TWOSHOT);
        else TFY_AR(__nv_bfloat16, TFY_ALGO_NVLS);
    } else {
        if (algo == TFY_ALGO_ONESHOT) TFY_AR(float, TFY_ALGO_ONESHOT);
        else if (algo == TFY_ALGO_TWOSHOT) TFY_AR(float, TFY_ALGO_TWOSHOT);
        else TFY_AR(float, TFY_ALGO_NVLS);
    }
#undef TFY_AR
    return (int)cudaGetLastError();
}

int tfy_broadcast(const TfyCommCtx* c, uint64_t off, size_t nbytes, int root, int use_nvls, int grid, int block,
                  cudaStream_t s) {
    if (nbytes % 16) return -2;
    const size_t n_packs = nbytes / 16;
    if (block <= 0) block = 512;
    if (grid <= 0) grid = tfy_pick_grid(n_packs, block * 2, 148);
    if (grid > TFY_MAX_BLOCKS) grid = TFY_MAX_BLOCKS;
    if (use_nvls && c->mc_base)
        tfy_broadcast_kernel<TFY_MODE_NVLS><<<grid, block, 0, s>>>(*c, off, n_packs, root);
    else
        tfy_broadcast_kernel<TFY_MODE_P2P><<<grid, block, 0, s>>>(*c, off, n_packs, root);
    return (int)cudaGetLastError();
}

int tfy_allgather(const TfyCommCtx* c, uint64_t off, size_t shard_bytes, int grid, int block, cudaStream_t s) {
    if (shard_bytes % 16) return -2;
    const size_t shard_packs = shard_bytes / 16;
    if (block <= 0) block = 512;
    if (grid <= 0) grid = tfy_pick_grid(shard_packs, block * 2, 148);
    if (grid > TFY_MAX_BLOCKS) grid = TFY_MAX_BLOCKS;
    tfy_allgather_kernel<<<grid, block, 0, s>>>(*c, off, shard_packs);
    return (int)cudaGetLastError();
}

// The fused gradient step.  shard_n: elements owned by each rank (multiple of 8).
// mode: 0 local (world==1), 1 P2P, 2 NVLS.
int tfy_fused_step_range(const TfyCommCtx* c, int grad_dtype, int param_dtype, int opt, int mode, uint64_t grad_off,
                         uint64_t param_off, size_t shard_n, float* master, float* s1, float* s2, TfyOptHyper* hp,
                         int zero_grads, int grid, int block, size_t e0, size_t e1, int advance, cudaStream_t s);

int tfy_fused_step(const TfyCommCtx* c, int grad_dtype, int param_dtype, int opt, int mode, uint64_t grad_off,
                   uint64_t param_off, size_t shard_n, float* master, float* s1, float* s2, TfyOptHyper* hp,
                   int zero_grads, int grid, int block, cudaStream_t s) {
    return tfy_fused_step_range(c, grad_dtype, param_dtype, opt, mode, grad_off, param_off, shard_n, master, s1, s2,
                                hp, zero_grads, grid, block, 0, (size_t)-1, 1, s);
}

// element range [e0, e1) of the flat buffers (multiples of 8; e1 is clamped to the buffer), see the kernel
int tfy_fused_step_range(const TfyCommCtx* c, int grad_dtype, int param_dtype, int opt, int mode, uint64_t grad_off,
                         uint64_t param_off, size_t shard_n, float* master, float* s1, float* s2, TfyOptHyper* hp,
                         int zero_grads, int grid, int block, size_t e0, size_t e1, int advance, cudaStream_t s) {
    if (shard_n % 8 || e0 % 8 || (e1 != (size_t)-1 && e1 % 8)) return -2;
    {
        const size_t total = shard_n * (size_t)(mode == TFY_MODE_LOCAL ? 1 : c->world);
        if (e1 > total) e1 = total;
        if (e0 > e1) e0 = e1;
    }
    if (mode == TFY_MODE_NVLS && c->mc_base == 0) return -5;
    if (block <= 0) block = (mode == TFY_MODE_LOCAL) ? 128 : 256;
    // single GPU: nothing to synchronise with, so oversubscribe the SMs for memory-level parallelism;
    // multi GPU: every CTA runs two cross-GPU barriers, keep one wave
    if (grid <= 0) {
        // work of the busiest rank: the largest overlap of [e0, e1) with one shard
        size_t span = e1 - e0;
        if (span > shard_n) span = shard_n;
        grid = tfy_pick_grid(span / 8 + 1, block, mode == TFY_MODE_LOCAL ? 148 * 6 : 148 * 2);
    }
    if (grid > TFY_MAX_BLOCKS) grid = TFY_MAX_BLOCKS;
#define TFY_FS4(GT, PT, O, M)                                                                                  \
    tfy_launch_pdl((tfy_fused_step_kernel<GT, PT, O, M>), dim3(grid), dim3(block), 0, s, *c, grad_off, param_off, shard_n, master, s1, s2, \
                                                                hp, zero_grads, e0, e1, advance)
#define TFY_FS3(GT, PT, O)                                   \
    do {                                                     \
        if (mode == TFY_MODE_LOCAL) TFY_FS4(GT, PT, O, TFY_MODE_LOCAL); \
        else if (mode == TFY_MODE_P2P) TFY_FS4(GT, PT, O, TFY_MODE_P2P); \
        else TFY_FS4(GT, PT, O, TFY_MODE_NVLS);              \
    } while (0)
#define TFY_FS2(GT, PT)                                            \
    do {                                                           \
        if (opt == TFY_OPT_SGD) TFY_FS3(GT, PT, TFY_OPT_SGD);      \
        else if (opt == TFY_OPT_ADADELTA) TFY_FS3(GT, PT, TFY_OPT_ADADELTA); \
        else if (opt == TFY_OPT_ADAM) TFY_FS3(GT, PT, TFY_OPT_ADAM); \
        else TFY_FS3(GT, PT, TFY_OPT_ADAGRAD);                     \
    } while (0)
    if (grad_dtype == TFY_BF16 && param_dtype == TFY_BF16) TFY_FS2(__nv_bfloat16, __nv_bfloat16);
    else if (grad_dtype == TFY_F32 && param_dtype == TFY_F32) TFY_FS2(float, float);
    else if (grad_dtype == TFY_BF16 && param_dtype == TFY_F32) TFY_FS2(__nv_bfloat16, float);
    else TFY_FS2(float, __nv_bfloat16);
#undef TFY_FS2
#undef TFY_FS3
#undef TFY_FS4
    return (int)cudaGetLastError();
}

}  // extern "C"
