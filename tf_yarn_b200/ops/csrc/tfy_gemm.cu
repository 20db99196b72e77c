// tcgen05 GEMM for sm_100a:  C[M,N] = act(A[M,K] . B[N,K]^T + bias[N])   (bf16 in, fp32 accumulate)
//
// This is the compute half of K5, the parameter-server PULL fused with the consuming GEMM:
// B (the layer's weight matrix) is addressed through a TMA tensor map whose base pointer may be a
// PEER address -- the ps rank's HBM mapped over NVLink by the symmetric arena -- so the weight
// tiles stream from the parameter server's memory straight into this SM's shared memory and into
// the tensor core, with no staging copy of the shard on the worker (reference path being replaced:
// TF gRPC RecvTensor of every variable before the first consumer op, tf_yarn/tensorflow/cluster.py:
// 60-66).  With a local B it is also the dense-layer GEMM of the mini-Keras fast path.
//
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d of a 128x64 A box and a 128x64 B box per
//              stage (128-byte swizzle), completion on an mbarrier (complete_tx)
//   warp 1   : TMEM allocation + MMA issuer -- one elected lane issues 4 x tcgen05.mma
//              (M128 N128 K16, kind::f16) per stage, tcgen05.commit releases the stage
//   warps 2-5: epilogue -- tcgen05.ld the fp32 accumulator out of TMEM (each warp owns the 32 TMEM
//              lanes of its quadrant), bias + activation, bf16 stores; or fp32 red.add for split-K
// Pipeline: STAGES-deep ring of {A,B} tiles guarded by full/empty mbarriers.
#include <cuda.h>

#include "tfy_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16, STAGES = 6;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;   // fp32 accumulator: 128 lanes x 128 columns
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp: the compiler keeps descriptors in uniform registers and emits back-to-back
// UTCHMMA / UTMALDG (with `lane == 0` each one was wrapped in an ELECT / R2UR / BRA.U.ANY loop, ~48 cycles).
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, 128-byte swizzled operand tile: rows of 64 bf16 (128 B); 8-row core groups 1024 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address        bits [0,14)
    d |= (uint64_t)0 << 16;                              // leading byte offset  bits [16,30): unused (1 atom on K)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset   bits [32,46): 8 rows x 128 B
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                              // layout: SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=128
__device__ __forceinline__ uint32_t make_idesc_bf16_m128(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace

// Fused split-K epilogue (out_mode 2): what runs after the partial sums of ALL K-slices of a tile have been added.
// The CTA whose arrival completes the tile reads the fp32 accumulator back (L2), clears it for the next step,
// applies bias + ReLU + dropout and writes the bf16 activations and the gradient-gate mask -- the work of the
// separate tfy_bias_act_drop_fwd_f32 launch of round 1 (same element indexing, same random numbers).
struct TfySplitKEpilogue {
    uint32_t* counters;      // two per output tile (arrive, finish), zero on entry, left zero
    __nv_bfloat16* y;        // [M, N] bf16 activations (row pitch ldy)
    uint8_t* mask;           // [M, N] bytes: 1 = gradient flows (optional)
    const TfyOptHyper* hp;   // dropout step counter
    float drop_rate;
    uint32_t seed;
    int ldy;
};

// out_mode 0: C (bf16) = act(acc + bias)      out_mode 1: C32 (fp32) += acc   (split-K partial sums)
// out_mode 2: as 1, then the last-arriving CTA of every tile runs the fused epilogue above
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tfy_gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     __nv_bfloat16* __restrict__ C, float* __restrict__ C32, const __nv_bfloat16* __restrict__ bias,
                     int M, int N, int K, int ldc, int relu, int out_mode, int k_tiles_per_split, TfySplitKEpilogue ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int total_k_tiles = (K + BK - 1) / BK;
    const int kt0 = blockIdx.z * k_tiles_per_split;
    int kt1 = kt0 + k_tiles_per_split;
    if (kt1 > total_k_tiles) kt1 = total_k_tiles;
    const int n_kt = kt1 - kt0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();                      // upstream grids complete; our dependents may start launching

    if (warp == 0) {
        if (tc_elect_one()) {
            for (int i = 0; i < n_kt; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* a_dst = tiles + (size_t)s * STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                const int k0 = (kt0 + i) * BK;
                tma_load_2d(&map_a, &full[s], a_dst, k0, m0);
                tma_load_2d(&map_b, &full[s], b_dst, k0, n0);
            }
        }
    } else if (warp == 1) {
        if (tc_elect_one()) {
            const uint32_t idesc = make_idesc_bf16_m128(BN);
            for (int i = 0; i < n_kt; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in 16-byte units
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                              (i > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit(&empty[s]);          // frees the smem stage when these MMAs have read it
            }
            tc_commit(tmem_full);              // accumulator complete
        }
    } else {
        // epilogue: this warp may only touch the TMEM lanes of its quadrant
        const int quad = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = m0 + quad * 32 + lane;
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, r);
            const int n = n0 + c;
            if (!row_ok || n >= N || n_kt <= 0) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (out_mode >= 1) {
                float* dst = C32 + (size_t)row * ldc + n;
                if (n + 16 <= N && (ldc & 3) == 0 && (n & 3) == 0) {
                    // 4 x red.v4.f32 instead of 16 scalar REDs
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]),
                                     "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                                     : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < N) atomicAdd(dst + j, v[j]);
                }
            } else {
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < N) v[j] += __bfloat162float(bias[n + j]);
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                __nv_bfloat16* dst = C + (size_t)row * ldc + n;
                if (n + 16 <= N && (ldc & 7) == 0 && (n & 7) == 0) {
                    tfy_st16(dst, TfyPack<__nv_bfloat16>::pack(v));
                    tfy_st16(dst + 8, TfyPack<__nv_bfloat16>::pack(v + 8));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < N) dst[j] = __float2bfloat16(v[j]);
                }
            }
        }
    }
    if (out_mode == 2 && warp >= 2) {
        // ---- fused split-K epilogue, spread over ALL K-slice CTAs of the tile ----
        // Every CTA arrives on the tile's counter after its red.adds, waits until all gridDim.z slices have arrived
        // (they are co-resident: the host only selects this mode when the whole grid fits the SMs) and then finishes
        // 1/gridDim.z of the tile: one 8-element group per thread, i.e. ONE L2 round trip.  (A first version let the
        // last-arriving CTA finish the whole 128x128 tile alone: 16 dependent L2 round trips per thread, +9 us.)
        uint32_t* cnt = ep.counters + 2 * (blockIdx.y * gridDim.x + blockIdx.x);
        asm volatile("bar.sync 1, 128;" ::: "memory");                 // this CTA's red.adds are all issued
        if (threadIdx.x == 64) {
            __threadfence();
            atomicAdd(cnt, 1u);
            while (*reinterpret_cast<volatile uint32_t*>(cnt) < gridDim.z) {
            }
            __threadfence();
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
            const uint32_t step = ep.hp ? (uint32_t)ep.hp->step : 0u;
            const float keep_scale = ep.drop_rate > 0.f ? 1.f / (1.f - ep.drop_rate) : 1.f;
            const int t = threadIdx.x - 64;                            // 0..127
            const int G = N / 8;                                       // 8-element groups per row (N % 8 == 0)
            constexpr int tile_g = BN / 8, tile_groups = BM * tile_g;
            const int per = (tile_groups + (int)gridDim.z - 1) / (int)gridDim.z;
            const int lo = (int)blockIdx.z * per, hi = min(lo + per, tile_groups);
            for (int i = lo + t; i < hi; i += 128) {
                const int r = i / tile_g, gq = i % tile_g;
                const int row = m0 + r, n = n0 + gq * 8;
                if (row >= M || n >= N) continue;
                float4* zp = reinterpret_cast<float4*>(C32 + (size_t)row * ldc + n);
                const float4 a = __ldcg(zp), b = __ldcg(zp + 1);
                zp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                zp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                if (bias) {
                    float bv[8];
                    TfyPack<__nv_bfloat16>::unpack(tfy_ld16(bias + n), bv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += bv[k];
                }
                const size_t idx = (size_t)row * G + n / 8;            // group index of the un-fused kernel
                uint32_t m_lo = 0, m_hi = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    bool on = true;
                    if (relu) { on = v[k] > 0.f; v[k] = on ? v[k] : 0.f; }
                    if (ep.drop_rate > 0.f) {
                        const bool keep = tfy_uniform(ep.seed, step, idx * 8 + k) >= ep.drop_rate;
                        v[k] = keep ? v[k] * keep_scale : 0.f;
                        on = on && keep;
                    }
                    if (k < 4) m_lo |= (on ? 1u : 0u) << (8 * k);
                    else m_hi |= (on ? 1u : 0u) << (8 * (k - 4));
                }
                tfy_st16(ep.y + (size_t)row * ep.ldy + n, TfyPack<__nv_bfloat16>::pack(v));
                if (ep.mask) *reinterpret_cast<uint2*>(ep.mask + (size_t)row * N + n) = make_uint2(m_lo, m_hi);
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {                                       // the CTA that finishes last re-arms both counters
            __threadfence();
            if (atomicAdd(cnt + 1, 1u) == gridDim.z - 1) {
                cnt[0] = 0u;
                cnt[1] = 0u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
bool g_attr_set = false;

bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess || !fn)
        return false;
    g_encode = reinterpret_cast<EncodeFn>(fn);
    return true;
}

// row-major [rows, K] bf16 matrix with leading dimension ld (elements); box = 64 (K) x 128 (rows)
bool make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld) {
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

// C[M,N] (bf16, ld ldc) = act(A[M,K] (ld lda) . B[N,K]^T (ld ldb) + bias[N]);  B may be a peer (NVLink) pointer.
// split_k > 1: partial products are accumulated with fp32 atomics into C32[M,N] (ld ldc, must be zeroed).
// Requirements: K % 8 == 0, lda % 8 == 0, ldb % 8 == 0, 16-byte aligned base pointers.
int tfy_gemm_bf16(const void* A, const void* B, void* C, float* C32, const void* bias, int M, int N, int K, int lda,
                  int ldb, int ldc, int relu, int split_k, cudaStream_t s) {
    if ((K & 7) || (lda & 7) || (ldb & 7)) return -2;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return -3;
    if (!load_encode()) return -4;
    if (!g_attr_set) {
        if (cudaFuncSetAttribute(tfy_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) !=
            cudaSuccess)
            return -5;
        g_attr_set = true;
    }
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, M, K, lda) || !make_map(&mb, B, N, K, ldb)) return -6;
    const int k_tiles = (K + BK - 1) / BK;
    if (split_k < 1) split_k = 1;
    if (split_k > k_tiles) split_k = k_tiles;
    const int per = (k_tiles + split_k - 1) / split_k;
    split_k = (k_tiles + per - 1) / per;
    const int out_mode = (split_k > 1 || (C == nullptr && C32 != nullptr)) ? 1 : 0;
    if (out_mode == 1 && C32 == nullptr) return -7;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, split_k);
    tfy_launch_pdl((tfy_gemm_bf16_kernel), dim3(grid), dim3(GEMM_THREADS), SMEM_BYTES, s, ma, mb, (__nv_bfloat16*)C, C32,
                                                                (const __nv_bfloat16*)bias, M, N, K, ldc, relu,
                                                                out_mode, per, TfySplitKEpilogue{});
    return (int)cudaGetLastError();
}

// Split-K GEMM with the fused epilogue: y[M,N] (bf16) = dropout(act(A . B^T + bias)), mask[M,N] (optional bytes).
// acc32 [M,N] fp32 (row pitch N) is scratch that must be zero on entry and is left zero; counters: two uint32 per
// 128x128 output tile, zero on entry, left zero.  N % 8 == 0.  The K-slice CTAs of a tile meet on a counter, so the
// whole grid must be co-resident: returns -8 when tiles x split_k exceeds the SM count (use the two-launch path).
int tfy_gemm_bf16_splitk_fused(const void* A, const void* B, float* acc32, uint32_t* counters, const void* bias, void* y,
                               void* mask, int M, int N, int K, int lda, int ldb, int relu, float drop_rate, uint32_t seed,
                               const TfyOptHyper* hp, int split_k, cudaStream_t s) {
    if ((K & 7) || (lda & 7) || (ldb & 7) || (N & 7)) return -2;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)y & 15)) return -3;
    if (!load_encode()) return -4;
    if (!g_attr_set) {
        if (cudaFuncSetAttribute(tfy_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) !=
            cudaSuccess)
            return -5;
        g_attr_set = true;
    }
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, M, K, lda) || !make_map(&mb, B, N, K, ldb)) return -6;
    const int k_tiles = (K + BK - 1) / BK;
    if (split_k < 1) split_k = 1;
    if (split_k > k_tiles) split_k = k_tiles;
    const int per = (k_tiles + split_k - 1) / split_k;
    split_k = (k_tiles + per - 1) / per;
    TfySplitKEpilogue ep;
    ep.counters = counters; ep.y = (__nv_bfloat16*)y; ep.mask = (uint8_t*)mask; ep.hp = hp;
    ep.drop_rate = drop_rate; ep.seed = seed; ep.ldy = N;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, split_k);
    {
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        if ((int)(grid.x * grid.y * grid.z) > sms) return -8;
    }
    tfy_launch_pdl((tfy_gemm_bf16_kernel), dim3(grid), dim3(GEMM_THREADS), SMEM_BYTES, s, ma, mb, (__nv_bfloat16*)nullptr,
                   acc32, (const __nv_bfloat16*)bias, M, N, K, N, relu, 2, per, ep);
    return (int)cudaGetLastError();
}

}  // extern "C"
