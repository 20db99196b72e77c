// Shared device helpers for the tf_yarn_b200 sm_100a kernels.
//
// Everything in ops/csrc is written for ONE target: Blackwell B200
// (-gencode arch=compute_100a,code=sm_100a).  There is no fallback path for
// other architectures.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TFY_MAX_RANKS 16
#define TFY_MAX_BLOCKS 1024            // barrier slots (one per CTA index)
#define TFY_FLAGS_BYTES (TFY_MAX_BLOCKS * TFY_MAX_RANKS * 4)

// Per-process view of the symmetric arena.  Passed BY VALUE to kernels (it is
// 160 bytes, well inside the 4 KB kernel parameter space) so that a CUDA graph
// captures the peer addresses together with the launch.
struct TfyCommCtx {
    uint64_t peer_base[TFY_MAX_RANKS];  // VA (in THIS process) of every rank's arena
    uint64_t mc_base;                   // multicast VA of the arena, 0 if NVLS is unavailable
    uint32_t* epoch;                    // local (non-shared) [TFY_MAX_BLOCKS][TFY_MAX_RANKS] counters
    int32_t rank;
    int32_t world;
};

// Device-resident optimizer hyper-parameters.  They live in HBM (not in kernel
// arguments) so that a captured CUDA graph sees learning-rate schedule changes
// and the step counter advancing without being re-captured.
struct TfyOptHyper {
    float lr;
    float p1;            // Adadelta: rho   | Adam: beta1 | SGD: momentum | Adagrad: initial acc (unused in kernel)
    float p2;            // Adam: beta2     | SGD: dampening
    float eps;
    float weight_decay;
    float grad_scale;    // multiplied into the (already averaged) gradient, e.g. 1/loss_scale
    int32_t step;        // completed optimizer steps
    int32_t flags;       // bit0: nesterov (SGD) / decoupled weight decay (Adam)
    uint32_t done;       // CTA completion counter used to advance `step` exactly once per launch
    uint32_t pad;
};

enum TfyDtype { TFY_BF16 = 0, TFY_F32 = 1 };
enum TfyAlgo { TFY_ALGO_ONESHOT = 0, TFY_ALGO_TWOSHOT = 1, TFY_ALGO_NVLS = 2 };
enum TfyOpt { TFY_OPT_SGD = 0, TFY_OPT_ADADELTA = 1, TFY_OPT_ADAM = 2, TFY_OPT_ADAGRAD = 3 };
enum TfyMode { TFY_MODE_LOCAL = 0, TFY_MODE_P2P = 1, TFY_MODE_NVLS = 2 };

// ---------------------------------------------------------------------------
// system-scope flag primitives (cross-GPU barrier on the signal pad)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tfy_red_release_sys_inc(uint32_t* addr) {
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(addr) : "memory");
}
__device__ __forceinline__ uint32_t tfy_ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// K8 — device-side barrier between the CTAs with the same blockIdx.x on every
// rank.  Monotonic counters: a rank adds 1 to flags[slot][me] on every peer
// (fire-and-forget `red.release.sys`, so the signal costs one one-way NVLink
// hop, not a CAS round trip) and then spins with `ld.acquire.sys` on its own
// pad until every peer's counter reached the local epoch.  The epoch lives in
// local HBM, so the protocol is replay-safe inside CUDA graphs.
__device__ __forceinline__ void tfy_block_barrier(const TfyCommCtx& c) {
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        const uint32_t slot = blockIdx.x;
        uint32_t* ep = c.epoch + slot * TFY_MAX_RANKS + peer;
        const uint32_t e = *ep + 1u;
        uint32_t* remote = reinterpret_cast<uint32_t*>(c.peer_base[peer]) + slot * TFY_MAX_RANKS + c.rank;
        tfy_red_release_sys_inc(remote);
        const uint32_t* mine =
            reinterpret_cast<const uint32_t*>(c.peer_base[c.rank]) + slot * TFY_MAX_RANKS + peer;
        while ((int32_t)(tfy_ld_acquire_sys(mine) - e) < 0) {
        }
        *ep = e;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// 16-byte vector helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 tfy_ld16(const void* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
// streaming (peer) load: do not allocate in L1, data is used once
__device__ __forceinline__ uint4 tfy_ld16_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void tfy_st16(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void tfy_st16_sys(void* p, uint4 v) {
    asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// NVLS: in-switch reduction of the 16 bytes at the same offset of every
// replica bound to the multicast object.  bf16 inputs are accumulated in fp32
// inside the switch (.acc::f32) and rounded once on the way out.
__device__ __forceinline__ uint4 tfy_mc_ld_reduce_bf16x8(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint4 tfy_mc_ld_reduce_f32x4(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
// NVLS: one store, replicated by the switch into every rank's HBM.
__device__ __forceinline__ void tfy_mc_st16(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float tfy_bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float tfy_bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t tfy_pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T>
struct TfyPack;  // 16-byte pack of T

template <>
struct TfyPack<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = tfy_bf16lo(v.x); f[1] = tfy_bf16hi(v.x);
        f[2] = tfy_bf16lo(v.y); f[3] = tfy_bf16hi(v.y);
        f[4] = tfy_bf16lo(v.z); f[5] = tfy_bf16hi(v.z);
        f[6] = tfy_bf16lo(v.w); f[7] = tfy_bf16hi(v.w);
    }
    __device__ static __forceinline__ void accum(const uint4& v, float* f) {
        f[0] += tfy_bf16lo(v.x); f[1] += tfy_bf16hi(v.x);
        f[2] += tfy_bf16lo(v.y); f[3] += tfy_bf16hi(v.y);
        f[4] += tfy_bf16lo(v.z); f[5] += tfy_bf16hi(v.z);
        f[6] += tfy_bf16lo(v.w); f[7] += tfy_bf16hi(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float* f) {
        uint4 v;
        v.x = tfy_pack_bf16x2(f[0], f[1]); v.y = tfy_pack_bf16x2(f[2], f[3]);
        v.z = tfy_pack_bf16x2(f[4], f[5]); v.w = tfy_pack_bf16x2(f[6], f[7]);
        return v;
    }
    __device__ static __forceinline__ uint4 mc_ld_reduce(const void* mc) { return tfy_mc_ld_reduce_bf16x8(mc); }
};

template <>
struct TfyPack<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ void accum(const uint4& v, float* f) {
        f[0] += __uint_as_float(v.x); f[1] += __uint_as_float(v.y);
        f[2] += __uint_as_float(v.z); f[3] += __uint_as_float(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float* f) {
        uint4 v;
        v.x = __float_as_uint(f[0]); v.y = __float_as_uint(f[1]);
        v.z = __float_as_uint(f[2]); v.w = __float_as_uint(f[3]);
        return v;
    }
    __device__ static __forceinline__ uint4 mc_ld_reduce(const void* mc) { return tfy_mc_ld_reduce_f32x4(mc); }
};


// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every hot-path kernel is launched with the programmatic stream
// serialisation attribute and starts with tfy_pdl_sync(): it lets ITS dependents start launching at once
// (all our grids are a single wave, so early dependents never take a slot a primary CTA still needs) and
// then waits for the grids it depends on to complete and flush.  Under stream capture these become
// programmatic edges of the CUDA graph: launch latency and each kernel's prologue (barrier init, TMEM
// allocation, tensor-map prefetch) overlap the tail of the previous kernel.  Opt-in (TFY_PDL=1): measured on
// the MNIST step it is a wash (fprop / first-layer wgrad -1 us each, dgrad +2.6 us with early dependents
// resident; 103.3 vs 101.5 us per step, profiles/bench_ours_N1_pdl_r1m.json), so the default stays off.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void tfy_pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

static inline bool tfy_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TFY_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t tfy_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tfy_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
