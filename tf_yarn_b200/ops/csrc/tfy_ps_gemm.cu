// K5, the parameter-server PULL fused with the GEMM that consumes it (sm_100a).
//
// First deep layer of a wide-and-deep model on the parameter-server path:
//     y[b, :] = concat_t( table_t[ids[t, b], :] ) ++ numeric[b, :]   .   W^T  + bias
// The reference's stack does this as T sparse RecvTensor pulls (one per embedding table), a concat and a MatMul
// against a weight that was itself pulled from another ps task (reference: tf_yarn/tensorflow/cluster.py:41-67).
// Here ONE kernel does it, and nothing it reads is ever copied to the worker first:
//   * A operand: producer warps GATHER the embedding rows of the batch straight from the ps ranks' HBM over NVLink
//     (16-byte peer loads of the fp32 master rows), convert them to bf16 and write them into the 128-byte-swizzled
//     shared-memory tile the tensor core reads -- the "embedding_bag" of a one-hot column is the operand producer
//     of the GEMM;
//   * B operand: the layer's weights stay on THEIR ps rank; TMA streams the bf16 shadow (row pitch padded to 8
//     elements) tile by tile into shared memory;
//   * tcgen05.mma M128 N256 K16 accumulates in TMEM; the epilogue adds the bias and stores bf16.
// The gathered activations are also written out once (bf16, [B, Kp]) because the backward needs them
// (dW = dy^T x, tfy_dense_bwd).  K chunk c < T is table c (embedding dim 64); chunk T holds the numeric features.
// warp 0: TMA (B tiles)   warp 1: TMEM + MMA issue   warps 2-17: A-tile producers (4 threads per row, 3 chunks in
// flight each: a gather is a dependent id -> row round trip to (remote) HBM per chunk, so the K loop is only as
// fast as the number of rows in flight)   warps 2-5 then run the epilogue
#include <cuda.h>

#include "tfy_common.cuh"

namespace {

constexpr int PG_BM = 128, PG_BN = 256, PG_BK = 64, PG_STAGES = 4;
constexpr int PG_A_BYTES = PG_BM * PG_BK * 2, PG_B_BYTES = PG_BN * PG_BK * 2, PG_STAGE_BYTES = PG_A_BYTES + PG_B_BYTES;
constexpr int PG_PROD_WARPS = 16, PG_PROD_THREADS = PG_PROD_WARPS * 32;   // 4 threads per A-tile row
constexpr int PG_DEPTH = 3;                                                  // gathered chunks in flight per thread
constexpr int PG_THREADS = (2 + PG_PROD_WARPS) * 32;                         // 576
constexpr size_t PG_SMEM = 1024 + (size_t)PG_STAGES * PG_STAGE_BYTES + 512;

__device__ __forceinline__ uint32_t pg_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pg_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pg_smem(bar)) : "memory");
}
__device__ __forceinline__ void pg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pg_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pg_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "PGWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra PGDONE;\n\t"
        "bra PGWAIT_LOOP;\n\t"
        "PGDONE:\n\t"
        "}" ::"r"(pg_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pg_tma_load(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(pg_smem(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(pg_smem(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ bool pg_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t pg_desc(uint32_t smem_addr) {      // K-major, SWIZZLE_128B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void pg_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void pg_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pg_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void pg_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float4 pg_ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

}  // namespace

// tables: device array of T peer pointers to fp32 [V, 64] embedding tables; ids: int64 [T, B]; numeric: fp32 [B, n_num]
// (n_num <= 64); map_w: bf16 shadow [N, Kp] of the weights (peer address allowed); xbuf: bf16 [B, Kp] (written);
// y: bf16 [B, N] = x . W^T + bias.
__global__ void __launch_bounds__(PG_THREADS, 1)
tfy_ps_gather_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const uint64_t* __restrict__ tables,
                          const long long* __restrict__ ids, const float* __restrict__ numeric,
                          const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ xbuf,
                          __nv_bfloat16* __restrict__ y, int B, int N, int T, int n_num, int Kp, long long V) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(ring + (size_t)PG_STAGES * PG_STAGE_BYTES);
    uint64_t* full_b = full_a + PG_STAGES;
    uint64_t* empty = full_b + PG_STAGES;
    uint64_t* acc_full = empty + PG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * PG_BM, n0 = blockIdx.y * PG_BN;
    const int n_chunks = T + (n_num > 0 ? 1 : 0);
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        for (int s = 0; s < PG_STAGES; ++s) {
            pg_mbar_init(&full_a[s], PG_PROD_THREADS);
            pg_mbar_init(&full_b[s], 1);
            pg_mbar_init(&empty[s], 1);
        }
        pg_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pg_smem(tmem_slot)),
                     "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();

    if (warp == 0) {
        // ===== weights: the ps rank's bf16 shadow, two 128-row boxes per 64-column chunk =====
        if (pg_elect()) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % PG_STAGES;
                if (c >= PG_STAGES) pg_mbar_wait(&empty[s], ((c / PG_STAGES) - 1) & 1);
                uint8_t* b_dst = ring + (size_t)s * PG_STAGE_BYTES + PG_A_BYTES;
                pg_mbar_expect_tx(&full_b[s], PG_B_BYTES);
                pg_tma_load(&map_w, &full_b[s], b_dst, c * PG_BK, n0);
                pg_tma_load(&map_w, &full_b[s], b_dst + PG_B_BYTES / 2, c * PG_BK, n0 + 128);
            }
        }
    } else if (warp == 1) {
        if (pg_elect()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PG_BN >> 3) << 17) |
                                   ((uint32_t)(PG_BM >> 4) << 24);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % PG_STAGES;
                const uint32_t ph = (c / PG_STAGES) & 1;
                pg_mbar_wait(&full_a[s], ph);
                pg_mbar_wait(&full_b[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = pg_smem(ring + (size_t)s * PG_STAGE_BYTES);
                const uint64_t adesc = pg_desc(a_addr), bdesc = pg_desc(a_addr + PG_A_BYTES);
#pragma unroll
                for (int k = 0; k < PG_BK / 16; ++k)
                    pg_umma(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (c > 0 || k > 0) ? 1u : 0u);
                pg_commit(&empty[s]);
            }
            pg_commit(acc_full);
        }
    } else {
        // ===== A-tile producers: 4 threads per sample; thread (r, p) owns columns [16 p, 16 p + 16) of every chunk =====
        const int pt = (int)threadIdx.x - 64;
        const int r = pt >> 2, p = pt & 3;
        const int b = b0 + r;
        const bool live = b < B;
        auto issue = [&](int c, float4* dst) {
            const float* row = nullptr;
            if (live && c < T) {
                const long long id = ids[(size_t)c * B + b];
                if (id >= 0 && id < V) row = reinterpret_cast<const float*>(tables[c]) + (size_t)id * 64 + p * 16;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = row ? pg_ld_peer(row + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 buf[PG_DEPTH][4];
#pragma unroll
        for (int j = 0; j < PG_DEPTH; ++j) issue(j, buf[j]);
        for (int c0 = 0; c0 < n_chunks; c0 += PG_DEPTH) {
#pragma unroll
            for (int j = 0; j < PG_DEPTH; ++j) {
                const int c = c0 + j;
                if (c < n_chunks) {
                    const int s = c % PG_STAGES;
                    uint4 pk[2];
                    if (c < T) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            pk[h].x = tfy_pack_bf16x2(buf[j][2 * h].x, buf[j][2 * h].y);
                            pk[h].y = tfy_pack_bf16x2(buf[j][2 * h].z, buf[j][2 * h].w);
                            pk[h].z = tfy_pack_bf16x2(buf[j][2 * h + 1].x, buf[j][2 * h + 1].y);
                            pk[h].w = tfy_pack_bf16x2(buf[j][2 * h + 1].z, buf[j][2 * h + 1].w);
                        }
                    } else {
                        float f[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            f[q] = (live && p * 16 + q < n_num) ? numeric[(size_t)b * n_num + p * 16 + q] : 0.f;
                        pk[0] = TfyPack<__nv_bfloat16>::pack(f);
                        pk[1] = TfyPack<__nv_bfloat16>::pack(f + 8);
                    }
                    if (c >= PG_STAGES) pg_mbar_wait(&empty[s], ((c / PG_STAGES) - 1) & 1);
                    uint8_t* a_row = ring + (size_t)s * PG_STAGE_BYTES + r * 128;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        *reinterpret_cast<uint4*>(a_row + (((2 * p + h) ^ (r & 7)) << 4)) = pk[h];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    pg_mbar_arrive(&full_a[s]);
                    if (live) {                          // keep the activations for the backward (dW = dy^T x)
                        __nv_bfloat16* xr = xbuf + (size_t)b * Kp + c * 64 + p * 16;
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (c * 64 + p * 16 + h * 8 + 8 <= Kp) tfy_st16(xr + h * 8, pk[h]);
                    }
                    issue(c + PG_DEPTH, buf[j]);         // refill this slot (zeros beyond the last table)
                }
            }
        }
        if (warp >= 6) goto pg_done;                     // warps 2-5 go on to the epilogue
        // ===== epilogue: bias, bf16, row-contiguous 16-byte stores =====
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                 // TMEM lane == accumulator row of this thread
        const int ob = b0 + row;
        pg_mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int cc = 0; cc < PG_BN; cc += 16) {
            uint32_t acc[16];
            pg_tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cc, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int n = n0 + cc;
            if (ob >= B || n >= N) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) + ((bias && n + j < N) ? __bfloat162float(bias[n + j]) : 0.f);
            __nv_bfloat16* dst = y + (size_t)ob * N + n;
            if (n + 16 <= N) {
                tfy_st16(dst, TfyPack<__nv_bfloat16>::pack(v));
                tfy_st16(dst + 8, TfyPack<__nv_bfloat16>::pack(v + 8));
            } else {
                for (int j = 0; j < 16 && n + j < N; ++j) dst[j] = __float2bfloat16(v[j]);
            }
        }
    }
pg_done:
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

namespace {
using PGEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PGEncodeFn pg_encode = nullptr;
bool pg_attr_set = false;
}  // namespace

extern "C" {

// Requirements: every table is fp32 [V, 64]; n_num <= 64; N % 8 == 0; Kp % 8 == 0 and Kp >= 64 T + n_num; the weight
// shadow is bf16 [N, Kp] (row pitch Kp), 16-byte aligned; xbuf bf16 [B, Kp], y bf16 [B, N].
int tfy_ps_gather_gemm(const void* w_shadow, const uint64_t* tables, const void* ids, const void* numeric,
                       const void* bias, void* xbuf, void* y, int B, int N, int T, int n_num, int Kp, long long V,
                       cudaStream_t s) {
    if (B < 1 || N < 8 || (N & 7) || (Kp & 7) || n_num < 0 || n_num > 64 || T < 0 || Kp < 64 * T + n_num) return -2;
    if (((uintptr_t)w_shadow | (uintptr_t)xbuf | (uintptr_t)y) & 15) return -3;
    if (!pg_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !fn)
            return -4;
        pg_encode = reinterpret_cast<PGEncodeFn>(fn);
    }
    if (!pg_attr_set) {
        if (cudaFuncSetAttribute(tfy_ps_gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PG_SMEM) !=
            cudaSuccess)
            return -5;
        pg_attr_set = true;
    }
    CUtensorMap mw;
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    if (pg_encode(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_shadow), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -6;
    dim3 grid((B + PG_BM - 1) / PG_BM, (N + PG_BN - 1) / PG_BN);
    tfy_launch_pdl((tfy_ps_gather_gemm_kernel), grid, dim3(PG_THREADS), PG_SMEM, s, mw, tables, (const long long*)ids,
                   (const float*)numeric, (const __nv_bfloat16*)bias, (__nv_bfloat16*)xbuf, (__nv_bfloat16*)y, B, N, T,
                   n_num, Kp, V);
    return (int)cudaGetLastError();
}

}  // extern "C"
