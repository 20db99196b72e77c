// tcgen05.mma operand-fetch probe for sm_100a (stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a).
//
// Two questions the guides do not answer and the conv kernels depend on:
//  (1) how fast is the tensor core fed from shared memory for each operand layout (un-swizzled
//      "interleave" core matrices vs 64/128-byte swizzles, K-major vs MN-major)?          -> `time` runs
//  (2) which physical 16-byte chunks does a swizzled descriptor read when its start address is not at
//      the start of the swizzle pattern (row-shifted start, stride between 8-row groups that is not
//      a multiple of the pattern, base_offset field)?                                      -> `probe` runs
// Every physical 16-byte chunk q of the A region holds (q & 255) in its even bf16 elements and (q >> 8)
// in the odd ones; B is a 16x16 identity, so D[m][k] = A[m][k] tells which chunk fed row m.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

struct Cfg {
    uint32_t m, n, a_mn, b_mn;
    uint32_t a_off, a_lbo, a_sbo, a_layout, a_base_offset;
    uint32_t b_off, b_lbo, b_sbo, b_layout;
    uint32_t reps, a_step, b_step, probe;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t bo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(bo & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, float* out, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t A_BYTES = 96 * 1024, B_OFF = 96 * 1024, B_BYTES = 64 * 1024;
    // A region: chunk-coded
    for (uint32_t q = threadIdx.x; q < A_BYTES / 16; q += blockDim.x) {
        __nv_bfloat16 lo = __float2bfloat16((float)(q & 255)), hi = __float2bfloat16((float)(q >> 8));
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)q * 16);
        for (int e = 0; e < 8; ++e) p[e] = (e & 1) ? hi : lo;
    }
    // B region: zeros, then (probe) a 16x16 identity, K-major un-swizzled: [k half][n][8 k], lbo 256, sbo 128
    for (uint32_t i = threadIdx.x; i < B_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + B_OFF)[i] = 0u;
    __syncthreads();
    if (c.probe && threadIdx.x < 16) {
        const int n = threadIdx.x, k = threadIdx.x;
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(smem + B_OFF + c.b_off + (k / 8) * 256 + n * 16);
        p[k % 8] = __float2bfloat16(1.0f);
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    uint32_t elected = 0;
    if (threadIdx.x < 32)
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    if (elected) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (c.a_mn << 15) | (c.b_mn << 16) |
                               ((c.n >> 3) << 17) | ((c.m >> 4) << 24);
        const uint32_t a0 = smem_u32(smem) + c.a_off, b0 = smem_u32(smem) + B_OFF + c.b_off;
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            ad[r] = mk_desc(a0 + r * c.a_step, c.a_lbo, c.a_sbo, c.a_layout, c.a_base_offset);
            bd[r] = mk_desc(b0 + r * c.b_step, c.b_lbo, c.b_sbo, c.b_layout, 0);
        }
        const long long t0 = clock64();
        if (c.reps == 1) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(ad[0]), "l"(bd[0]), "r"(idesc), "r"(0u)
                : "memory");
        } else {
#pragma unroll 1
            for (uint32_t r = 0; r < c.reps; r += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                        "l"(ad[u]), "l"(bd[u]), "r"(idesc), "r"(1u)
                        : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
        asm volatile(
            "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::
                "r"(smem_u32(&bar))
            : "memory");
        cycles[0] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int warp = threadIdx.x >> 5;
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

struct Named { const char* name; Cfg c; };

int main() {
    const size_t smem = 96 * 1024 + 64 * 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* out;
    long long* cyc;
    cudaMalloc(&out, 128 * 16 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    const uint32_t NONE = 0, SW128 = 2, SW64 = 4, R = 512;
    std::vector<Named> runs = {
        // ---- timing: {m, n, a_mn, b_mn, a_off, a_lbo, a_sbo, a_layout, a_bo, b_off, b_lbo, b_sbo, b_layout, reps, a_step, b_step, probe}
        {"t_sw128_K_m128n64", {128, 64, 0, 0, 0, 0, 1024, SW128, 0, 0, 0, 1024, SW128, R, 32, 32, 0}},
        {"t_sw64_K_m128n64", {128, 64, 0, 0, 0, 0, 512, SW64, 0, 0, 0, 512, SW64, R, 32, 32, 0}},
        {"t_none_K_dense_m128n64", {128, 64, 0, 0, 0, 128, 256, NONE, 0, 0, 128, 256, NONE, R, 4096, 2048, 0}},
        {"t_none_K_fprop_m128n64", {128, 64, 0, 0, 0, 3216, 160, NONE, 0, 0, 1040, 128, NONE, R, 320, 4160, 0}},
        {"t_none_K_fprop_shift16", {128, 64, 0, 0, 16, 3216, 160, NONE, 0, 0, 1040, 128, NONE, R, 320, 4160, 0}},
        {"t_none_K_rows128_m128n64", {128, 64, 0, 0, 0, 2048, 128, NONE, 0, 0, 1040, 128, NONE, R, 0, 0, 0}},
        {"t_none_wgrad_MNMN_m64n32", {64, 32, 1, 1, 0, 128, 2064, NONE, 0, 0, 160, 3216, NONE, R, 256, 320, 0}},
        {"t_swz_wgrad_MNMN_m64n32", {64, 32, 1, 1, 0, 0, 1024, SW128, 0, 0, 0, 512, SW64, R, 2048, 1024, 0}},
        {"t_none_dgrad_K_MN_m128n32", {128, 32, 0, 1, 0, 3216, 160, NONE, 0, 0, 128, 1040, NONE, R, 320, 256, 0}},
        {"t_swz_dgrad_K_MN_m128n32", {128, 32, 0, 1, 0, 0, 1024, SW128, 0, 0, 0, 512, SW64, R, 32, 1024, 0}},
        {"t_sw128_K_m128n32", {128, 32, 0, 0, 0, 0, 1024, SW128, 0, 0, 0, 1024, SW128, R, 32, 32, 0}},
        {"t_sw128_K_m64n32", {64, 32, 0, 0, 0, 0, 1024, SW128, 0, 0, 0, 1024, SW128, R, 32, 32, 0}},
        {"t_sw128_K_m128n128", {128, 128, 0, 0, 0, 0, 1024, SW128, 0, 0, 0, 1024, SW128, R, 32, 32, 0}},
        {"t_sw128_K_shiftrow_sbo1280", {128, 64, 0, 0, 128, 0, 1280, SW128, 1, 0, 0, 1024, SW128, R, 32, 32, 0}},
        // ---- probes (B = identity 16x16, un-swizzled K-major: lbo 256, sbo 128)
        {"p_none_K_lbo3216_sbo160_off16", {128, 16, 0, 0, 16, 3216, 160, NONE, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_aligned", {128, 16, 0, 0, 0, 0, 1024, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_off128_bo0", {128, 16, 0, 0, 128, 0, 1024, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_off128_bo1", {128, 16, 0, 0, 128, 0, 1024, SW128, 1, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_sbo1280_bo0", {128, 16, 0, 0, 0, 0, 1280, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_off256_sbo1280_bo2", {128, 16, 0, 0, 256, 0, 1280, SW128, 2, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_K_off32_kadvance", {128, 16, 0, 0, 32, 0, 1024, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw64_K_aligned", {128, 16, 0, 0, 0, 0, 512, SW64, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw64_K_off64_bo0", {128, 16, 0, 0, 64, 0, 512, SW64, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw64_K_sbo640_bo0", {128, 16, 0, 0, 0, 0, 640, SW64, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_MN_m64", {64, 16, 1, 0, 0, 0, 1024, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
        {"p_sw128_MN_m64_off128", {64, 16, 1, 0, 128, 0, 1024, SW128, 0, 0, 256, 128, NONE, 1, 0, 0, 1}},
    };
    std::vector<float> h(128 * 16);
    for (auto& r : runs) {
        cudaMemset(out, 0, 128 * 16 * sizeof(float));
        probe_kernel<<<1, 128, smem>>>(r.c, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long cy = 0;
        cudaMemcpy(&cy, cyc, sizeof(cy), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            printf("{\"name\": \"%s\", \"error\": \"%s\"}\n", r.name, cudaGetErrorString(e));
            return 1;
        }
        if (!r.c.probe) {
            printf("{\"name\": \"%s\", \"cycles_per_mma\": %.1f}\n", r.name, (double)cy / r.c.reps);
        } else {
            cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
            // first chunk: (D[m][0], D[m][1]) -> q ; second chunk: (D[m][8], D[m][9])
            printf("{\"name\": \"%s\", \"q0\": [", r.name);
            const int rows = r.c.m == 64 ? 64 : 128;
            for (int m = 0; m < rows; ++m) {
                // M = 64 accumulators live in lanes (m % 16) + 32 * (m / 16)
                const int lane = r.c.m == 64 ? (m % 16) + 32 * (m / 16) : m;
                int q = r.c.a_mn ? (int)h[lane * 16 + 0] : (int)h[lane * 16 + 0] + 256 * (int)h[lane * 16 + 1];
                printf("%d%s", q, m + 1 < rows ? "," : "");
            }
            printf("], \"q1\": [");
            for (int m = 0; m < rows; ++m) {
                const int lane = r.c.m == 64 ? (m % 16) + 32 * (m / 16) : m;
                int q = r.c.a_mn ? (int)h[lane * 16 + 1] : (int)h[lane * 16 + 8] + 256 * (int)h[lane * 16 + 9];
                printf("%d%s", q, m + 1 < rows ? "," : "");
            }
            printf("]}\n");
        }
    }
    return 0;
}
