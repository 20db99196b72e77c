// tcgen05 implicit-GEMM 3x3 convolution kernels (C_in = 32 -> C_out = 64, stride 1, VALID) for sm_100a:
// the second convolution of the MNIST-CNN, forward (fused with bias + ReLU + 2x2 max-pool + dropout),
// data gradient (fused with the ReLU gate of the previous layer) and weight gradient.
//
// Shared-memory operand format (all three kernels).  A CTA works on an 8 x 8 pixel patch of an image
// PAIR.  TMA (5-D tensor map over the NHWC tensor, channel dimension split into groups of 8) stores a
// patch as
//        [channel group][h][image n][w][8 channels = 16 B]
// i.e. the un-swizzled ("interleave") UMMA canonical layout: a core matrix is 8 consecutive w (8 x 16 B
// = 128 contiguous bytes), the next core matrix along the pixel dimension is the next (h, n) row, and
// the next one along the channel dimension is the next channel group.  Because the layout is not
// swizzled, a filter tap (kh, kw) is just a different START ADDRESS into the same halo patch
// (+kh rows, +kw pixels): the halo is loaded ONCE and the nine im2col views never materialise, neither
// in HBM nor in shared memory.  The pixel dimension is GEMM-M for fprop / dgrad (K-major A operand) and
// GEMM-K for wgrad (MN-major operands: the SAME bytes described with the other major-ness).
//
//   fprop : D[pix, o]  = sum_{t,c} a[pix + t, c] W[o, t, c]     M=128 N=64 K=9x32   (A: K-major, B: K-major)
//   dgrad : D[pix, c]  = sum_{t,o} dz[pix - t, o] W[o, t, c]    M=128 N=32 K=9x64   (A: K-major, B: MN-major)
//   wgrad : D[o, (t,c)] = sum_pix dz[pix, o] a[pix + t, c]      M=64  N=32 K=128/patch, 9 accumulators
//                                                               (A: MN-major, B: MN-major), persistent CTAs
// Weights ([O][3][3][C] bf16) are loaded by ONE TMA as [tap][c group][o][8 c]: K-major for fprop and,
// read with the MN-major flag, the transposed operand dgrad needs -- no transpose kernel.
#include <cuda.h>

#include "tfy_common.cuh"

namespace {

constexpr int CIN = 32, COUT = 64, TAPS = 9;
constexpr int CG_IN = CIN / 8, CG_OUT = COUT / 8;
constexpr int HALO = 10;                               // 8 + 2
constexpr int ROW_B = HALO * 16;                       // 160  : one (h, n) row of a halo patch
constexpr int HROW_B = 2 * ROW_B;                      // 320  : one h step (2 images)
constexpr int HALO_G_B = HALO * HROW_B;                // 3200 : one channel group of a halo patch
constexpr int PATCH_G_B = 8 * 2 * 8 * 16;              // 2048 : one channel group of an 8x8x2 patch
constexpr int W_TAP_B = CG_IN * COUT * 16;             // 4096
constexpr int W_BYTES = TAPS * W_TAP_B;                // 36864
constexpr int CONV_THREADS = 192;

constexpr int FPROP_A_B = CG_IN * HALO_G_B;            // 12800
constexpr int DGRAD_A_B = CG_OUT * HALO_G_B;           // 25600
constexpr size_t FPROP_SMEM = FPROP_A_B + W_BYTES + 128 + 128;
constexpr size_t DGRAD_SMEM = DGRAD_A_B + W_BYTES + 128 + 128;
constexpr int WG_STAGES = 4;
constexpr int WG_DZ_B = CG_OUT * PATCH_G_B;            // 16384
constexpr int WG_STAGE_B = WG_DZ_B + FPROP_A_B;        // 29184
constexpr size_t WGRAD_SMEM = (size_t)WG_STAGES * WG_STAGE_B + 128 + 256;

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void c_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void c_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "CWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra CDONE;\n\t"
        "bra CWAIT_LOOP;\n\t"
        "CDONE:\n\t"
        "}" ::"r"(c_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void c_tma_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                         int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(c_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void c_tma_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                         int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(c_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
// Un-swizzled UMMA shared-memory descriptor.  K-major operand: lbo = byte distance between the two
// 8-element halves of the K=16 slice, sbo = distance between consecutive groups of 8 rows.  MN-major
// operand: lbo = distance between consecutive groups of 8 along K, sbo = between groups of 8 along M/N.
__device__ __forceinline__ uint64_t c_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16
__device__ __forceinline__ uint32_t c_idesc(uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
           ((m >> 4) << 24);
}
__device__ __forceinline__ void c_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void c_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void c_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void c_tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c_smem_u32(slot)),
                 "r"((uint32_t)COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void c_tmem_free(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"((uint32_t)COLS) : "memory");
}
__device__ __forceinline__ void c_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t c_hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// must stay identical to tfy_uniform() of tfy_nn.cu (the backward kernels only see the code byte)
__device__ __forceinline__ float c_uniform(uint32_t seed, uint32_t step, uint64_t idx) {
    uint32_t h = c_hash32(seed ^ c_hash32(step * 0x9E3779B9U + 0x85ebca6bU) ^
                          c_hash32((uint32_t)idx * 0xC2B2AE35U + (uint32_t)(idx >> 32) + 0x27d4eb2fU));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ uint8_t* c_align128(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// forward: pooled = dropout(maxpool2x2(relu(conv(a, W) + bias))), code byte per pooled element
// (bits 0-1 = argmax position dy*2+dx inside the 2x2 window, bit 2 = gradient flows).
// grid = (OW/8, OH/8, B/2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS, 4)
tfy_conv3x3_fprop_pool_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                              const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ pooled,
                              uint8_t* __restrict__ code, int OH, int OW, float drop_rate, uint32_t seed,
                              const TfyOptHyper* __restrict__ hp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* a_tile = c_align128(smem_raw);                   // [4][10][2][10][16 B]
    uint8_t* w_tile = a_tile + FPROP_A_B;                     // [9][4][64][16 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);   // 0: operands landed, 1: accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ow0 = blockIdx.x * 8, oh0 = blockIdx.y * 8, b0 = blockIdx.z * 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        c_mbar_init(&bars[0], 1);
        c_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) c_tmem_alloc<64>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            c_mbar_expect_tx(&bars[0], FPROP_A_B + W_BYTES);
            c_tma_5d(&map_a, &bars[0], a_tile, 0, ow0, b0, oh0, 0);
            c_tma_4d(&map_w, &bars[0], w_tile, 0, 0, 0, 0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = c_idesc(128, COUT, 0, 0);
            const uint32_t a0 = c_smem_u32(a_tile), w0 = c_smem_u32(w_tile);
            c_mbar_wait(&bars[0], 0);
            c_fence_after();
#pragma unroll 1
            for (int t = 0; t < TAPS; ++t) {
                const int kh = t / 3, kw = t % 3;
#pragma unroll
                for (int ks = 0; ks < CIN / 16; ++ks) {
                    // A: rows = pixels (8 w per core matrix, (h,n) groups ROW_B apart), K halves = channel groups
                    const uint64_t adesc = c_desc(a0 + kh * HROW_B + kw * 16 + ks * 2 * HALO_G_B, HALO_G_B, ROW_B);
                    // B: rows = output channels (16 B apart, groups of 8 = 128 B), K halves = channel groups
                    const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 2 * (COUT * 16), COUT * 16, 128);
                    c_umma(tmem_base, adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                }
            }
            c_commit(&bars[1]);
        }
    } else {
        // epilogue warps 2..5 own TMEM lane quadrants 2,3,0,1.  Accumulator row r = (h, n, w) = (r/16, (r/8)&1, r%8)
        const int quad = warp & 3;
        c_mbar_wait(&bars[1], 0);
        c_fence_after();
        const int r = quad * 32 + lane;
        const int h = r >> 4, n = (r >> 3) & 1, w = r & 7;
        const int PH = OH / 2, PW = OW / 2;
        const uint32_t step = hp ? (uint32_t)hp->step : 0u;
        const float keep_scale = drop_rate > 0.f ? 1.f / (1.f - drop_rate) : 1.f;
        const bool writer = ((lane & 16) == 0) && ((lane & 1) == 0);          // h even, w even
        const size_t prow = (((size_t)(b0 + n) * PH + (oh0 + h) / 2) * PW + (ow0 + w) / 2) * COUT;
#pragma unroll 1
        for (int c = 0; c < COUT; c += 16) {
            uint32_t acc[16];
            c_tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, acc);
            float outv[16];
            uint32_t codes[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(acc[j]);
                // window of the (even h, even w) lane: itself (0,0), lane^1 (0,1), lane^16 (1,0), lane^17 (1,1)
                const float v01 = __shfl_xor_sync(0xffffffffu, v, 1);
                const float v10 = __shfl_xor_sync(0xffffffffu, v, 16);
                const float v11 = __shfl_xor_sync(0xffffffffu, v, 17);
                float best = v;
                int arg = 0;
                if (v01 > best) { best = v01; arg = 1; }
                if (v10 > best) { best = v10; arg = 2; }
                if (v11 > best) { best = v11; arg = 3; }
                float o = best + __bfloat162float(bias[c + j]);
                bool on = o > 0.f;
                o = on ? o : 0.f;
                if (drop_rate > 0.f) {
                    const bool keep = c_uniform(seed, step, prow + c + j) >= drop_rate;
                    o = keep ? o * keep_scale : 0.f;
                    on = on && keep;
                }
                outv[j] = o;
                codes[j >> 2] |= ((uint32_t)arg | (on ? 4u : 0u)) << (8 * (j & 3));
            }
            if (writer) {
                tfy_st16(pooled + prow + c, TfyPack<__nv_bfloat16>::pack(outv));
                tfy_st16(pooled + prow + c + 8, TfyPack<__nv_bfloat16>::pack(outv + 8));
                *reinterpret_cast<uint4*>(code + prow + c) = make_uint4(codes[0], codes[1], codes[2], codes[3]);
            }
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == 1) c_tmem_free<64>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// data gradient: dx[b, y, x, c] = sum_{kh,kw,o} dz[b, y-kh, x-kw, o] W[o, kh, kw, c], optionally gated by
// the previous layer's ReLU (gate = that layer's output: dx is zeroed where gate <= 0).
// dz: [B, H-2, W-2, 64]; dx: [B, H, W, 32].  grid = (ceil(W/8), ceil(H/8), B/2); TMA zero-fills the halo.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS, 3)
tfy_conv3x3_dgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_w,
                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ dx, int H, int W) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* a_tile = c_align128(smem_raw);                   // [8][10][2][10][16 B]
    uint8_t* w_tile = a_tile + DGRAD_A_B;                     // [9][4][64][16 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * 8, y0 = blockIdx.y * 8, b0 = blockIdx.z * 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_dz)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        c_mbar_init(&bars[0], 1);
        c_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) c_tmem_alloc<32>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            c_mbar_expect_tx(&bars[0], DGRAD_A_B + W_BYTES);
            c_tma_5d(&map_dz, &bars[0], a_tile, 0, x0 - 2, b0, y0 - 2, 0);     // out-of-range pixels read as zero
            c_tma_4d(&map_w, &bars[0], w_tile, 0, 0, 0, 0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = c_idesc(128, CIN, 0, 1);                    // B is MN-major: N = c contiguous
            const uint32_t a0 = c_smem_u32(a_tile), w0 = c_smem_u32(w_tile);
            c_mbar_wait(&bars[0], 0);
            c_fence_after();
#pragma unroll 1
            for (int t = 0; t < TAPS; ++t) {
                const int kh = t / 3, kw = t % 3;
#pragma unroll
                for (int ks = 0; ks < COUT / 16; ++ks) {
                    // halo origin is (y0-2, x0-2): dz[y-kh, x-kw] sits (2-kh, 2-kw) into the patch
                    const uint64_t adesc =
                        c_desc(a0 + (2 - kh) * HROW_B + (2 - kw) * 16 + ks * 2 * HALO_G_B, HALO_G_B, ROW_B);
                    // W tap as [c group][o][8 c]: K = o (16 B apart, groups of 8 o = 128 B), N groups = c groups
                    const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 16 * 16, 128, COUT * 16);
                    c_umma(tmem_base, adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                }
            }
            c_commit(&bars[1]);
        }
    } else {
        const int quad = warp & 3;
        c_mbar_wait(&bars[1], 0);
        c_fence_after();
        const int r = quad * 32 + lane;
        const int h = r >> 4, n = (r >> 3) & 1, w = r & 7;
        const int y = y0 + h, x = x0 + w;
        const bool valid = y < H && x < W;
        const size_t off = (((size_t)(b0 + n) * H + y) * W + x) * CIN;
#pragma unroll
        for (int c = 0; c < CIN; c += 16) {
            uint32_t acc[16];
            c_tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, acc);
            if (valid) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
                if (gate) {
                    float g[16];
                    TfyPack<__nv_bfloat16>::unpack(tfy_ld16(gate + off + c), g);
                    TfyPack<__nv_bfloat16>::unpack(tfy_ld16(gate + off + c + 8), g + 8);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = g[j] > 0.f ? v[j] : 0.f;
                }
                tfy_st16(dx + off + c, TfyPack<__nv_bfloat16>::pack(v));
                tfy_st16(dx + off + c + 8, TfyPack<__nv_bfloat16>::pack(v + 8));
            }
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == 1) c_tmem_free<32>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[o, t, c] = sum_{b,y,x} dz[b, y, x, o] a[b, y+kh, x+kw, c].  Persistent CTAs walk the
// 8x8x2 pixel patches (the GEMM K dimension) through a TMA ring, accumulating nine 64x32 tiles in TMEM
// (one per tap: 288 of the 512 columns); partial sums meet in a zeroed fp32 buffer (red.global.add.v4.f32)
// and, after a grid-wide arrive/release, every CTA converts and re-zeroes its slice of it.
// acc32: [64*288] floats, zero on entry (left zero on exit); sync: two uint32, zero-initialised once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS, 1)
tfy_conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_a,
                         float* __restrict__ acc32, __nv_bfloat16* __restrict__ dw, uint32_t* __restrict__ sync,
                         int tiles_x, int tiles_y, int n_patches) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = c_align128(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)WG_STAGES * WG_STAGE_B);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tmem_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gen0 = threadIdx.x == 0 ? *reinterpret_cast<volatile uint32_t*>(sync + 1) : 0u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_dz)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        for (int s = 0; s < WG_STAGES; ++s) { c_mbar_init(&full[s], 1); c_mbar_init(&empty[s], 1); }
        c_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) c_tmem_alloc<512>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % WG_STAGES;
                if (i >= WG_STAGES) c_mbar_wait(&empty[s], ((i / WG_STAGES) - 1) & 1);
                const int p = (int)blockIdx.x + i * (int)gridDim.x;
                const int tx = p % tiles_x, ty = (p / tiles_x) % tiles_y, bz = p / (tiles_x * tiles_y);
                uint8_t* st = stages + (size_t)s * WG_STAGE_B;
                c_mbar_expect_tx(&full[s], WG_STAGE_B);
                c_tma_5d(&map_dz, &full[s], st, 0, tx * 8, bz * 2, ty * 8, 0);            // [8 og][8 h][2 n][8 w][8 o]
                c_tma_5d(&map_a, &full[s], st + WG_DZ_B, 0, tx * 8, bz * 2, ty * 8, 0);   // [4 cg][10 h][2 n][10 w][8 c]
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = c_idesc(64, CIN, 1, 1);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % WG_STAGES;
                c_mbar_wait(&full[s], (i / WG_STAGES) & 1);
                c_fence_after();
                const uint32_t dz0 = c_smem_u32(stages + (size_t)s * WG_STAGE_B), a0 = dz0 + WG_DZ_B;
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t % 3;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {          // 16 pixels (one h row of both images) per MMA
                        // A = dz^T: M = o (8 per 16 B; o groups PATCH_G_B apart), K = pixels (8 w = 128 B per group)
                        const uint64_t adesc = c_desc(dz0 + j * 256, 128, PATCH_G_B);
                        // B = a (tap shifted): N = c (c groups HALO_G_B apart), K = pixels ((h, n) rows ROW_B apart)
                        const uint64_t bdesc = c_desc(a0 + (j + kh) * HROW_B + kw * 16, ROW_B, HALO_G_B);
                        c_umma(tmem_base + (uint32_t)(t * CIN), adesc, bdesc, idesc, (i > 0 || j > 0) ? 1u : 0u);
                    }
                }
                c_commit(&empty[s]);                        // smem stage reusable once these MMAs retire
            }
            c_commit(tmem_full);
        }
    } else if (my_patches > 0) {
        // M = 64 accumulator: row o lives in TMEM lane (o % 16) + 32 * (o / 16): lanes 0..15 of each quadrant
        const int quad = warp & 3;
        c_mbar_wait(tmem_full, 0);
        c_fence_after();
        const int o = quad * 16 + lane;
#pragma unroll 1
        for (int col = 0; col < TAPS * CIN; col += 16) {
            uint32_t acc[16];
            c_tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col, acc);
            if (lane < 16) {
                float* dst = acc32 + (size_t)o * (TAPS * CIN) + col;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                 "f"(__uint_as_float(acc[j])), "f"(__uint_as_float(acc[j + 1])),
                                 "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3]))
                                 : "memory");
            }
        }
        __threadfence();
    }
    c_fence_before();
    __syncthreads();
    if (warp == 1) c_tmem_free<512>(tmem_base);

    // grid-wide arrive / release (all CTAs are co-resident: grid <= #SMs, one CTA per SM)
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t arrived = atomicAdd(sync, 1u);
        if (arrived == gridDim.x - 1) {
            *reinterpret_cast<volatile uint32_t*>(sync) = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        } else {
            while (*reinterpret_cast<volatile uint32_t*>(sync + 1) == gen0) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
    constexpr int TOTAL4 = COUT * TAPS * CIN / 4;
    for (int i = (int)blockIdx.x * CONV_THREADS + (int)threadIdx.x; i < TOTAL4; i += (int)gridDim.x * CONV_THREADS) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(acc32) + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 packed;
        packed.x = *reinterpret_cast<uint32_t*>(&lo);
        packed.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dw + (size_t)i * 4) = packed;
        __stcg(reinterpret_cast<float4*>(acc32) + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

namespace {
using CEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
CEncodeFn c_encode = nullptr;
bool c_attr_set = false;

bool c_init() {
    if (!c_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !fn)
            return false;
        c_encode = reinterpret_cast<CEncodeFn>(fn);
    }
    if (!c_attr_set) {
        if (cudaFuncSetAttribute(tfy_conv3x3_fprop_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)FPROP_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)DGRAD_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)WGRAD_SMEM) != cudaSuccess)
            return false;
        c_attr_set = true;
    }
    return true;
}

// NHWC bf16 tensor [B, H, W, C] seen as {8 c, W, B, H, C/8}; box = {8, bw, 2, bh, C/8}
bool c_map_nhwc(CUtensorMap* m, const void* p, int B, int H, int W, int C, int bw, int bh) {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H, (cuuint64_t)(C / 8)};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2, 16};
    cuuint32_t box[5] = {8, (cuuint32_t)bw, 2, (cuuint32_t)bh, (cuuint32_t)(C / 8)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return c_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weights [O][9][C] bf16 seen as {8 c, O, C/8, 9}; one box = everything, stored [tap][c group][o][8 c]
bool c_map_w(CUtensorMap* m, const void* p) {
    cuuint64_t dims[4] = {8, (cuuint64_t)COUT, (cuuint64_t)CG_IN, (cuuint64_t)TAPS};
    cuuint64_t strides[3] = {(cuuint64_t)TAPS * CIN * 2, 16, (cuuint64_t)CIN * 2};
    cuuint32_t box[4] = {8, (cuuint32_t)COUT, (cuuint32_t)CG_IN, (cuuint32_t)TAPS};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return c_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" {

// a: [B, H, W, 32] bf16 (NHWC), w: [64, 3, 3, 32] bf16, bias: [64] bf16
// pooled / code: [B, (H-2)/2, (W-2)/2, 64].  Requires (H-2) % 8 == 0, (W-2) % 8 == 0, B % 2 == 0.
int tfy_conv3x3_c32_pool_fwd(const void* a, const void* w, const void* bias, void* pooled, void* code, int B, int H,
                             int W, float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2)) return -2;
    if (!c_init()) return -4;
    CUtensorMap ma, mw;
    if (!c_map_nhwc(&ma, a, B, H, W, CIN, HALO, HALO) || !c_map_w(&mw, w)) return -6;
    dim3 grid(OW / 8, OH / 8, B / 2);
    tfy_conv3x3_fprop_pool_kernel<<<grid, CONV_THREADS, FPROP_SMEM, s>>>(
        ma, mw, (const __nv_bfloat16*)bias, (__nv_bfloat16*)pooled, (uint8_t*)code, OH, OW, drop_rate, seed, hp);
    return (int)cudaGetLastError();
}

// dz: [B, H-2, W-2, 64] bf16, w: [64, 3, 3, 32], gate (optional): [B, H, W, 32], dx: [B, H, W, 32].  B % 2 == 0.
int tfy_conv3x3_c32_dgrad(const void* dz, const void* w, const void* gate, void* dx, int B, int H, int W,
                          cudaStream_t s) {
    if (B % 2) return -2;
    if (!c_init()) return -4;
    CUtensorMap mz, mw;
    if (!c_map_nhwc(&mz, dz, B, H - 2, W - 2, COUT, HALO, HALO) || !c_map_w(&mw, w)) return -6;
    dim3 grid((W + 7) / 8, (H + 7) / 8, B / 2);
    tfy_conv3x3_dgrad_kernel<<<grid, CONV_THREADS, DGRAD_SMEM, s>>>(mz, mw, (const __nv_bfloat16*)gate,
                                                                    (__nv_bfloat16*)dx, H, W);
    return (int)cudaGetLastError();
}

// a: [B, H, W, 32], dz: [B, H-2, W-2, 64], dw: [64, 3, 3, 32] bf16 (overwritten).
// acc32: 64*288 floats, zero on entry and on exit; sync: 2 x uint32, zeroed once at allocation.
int tfy_conv3x3_c32_wgrad(const void* a, const void* dz, float* acc32, void* dw, uint32_t* sync, int B, int H, int W,
                          cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2)) return -2;
    if (!c_init()) return -4;
    CUtensorMap mz, ma;
    if (!c_map_nhwc(&mz, dz, B, OH, OW, COUT, 8, 8) || !c_map_nhwc(&ma, a, B, H, W, CIN, HALO, HALO)) return -6;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    int grid = n_patches < 148 ? n_patches : 148;
    // an even split keeps every CTA on the same number of patches (576 patches -> 144 CTAs x 4)
    for (int g = grid; g >= 96; --g)
        if (n_patches % g == 0) { grid = g; break; }
    tfy_conv3x3_wgrad_kernel<<<grid, CONV_THREADS, WGRAD_SMEM, s>>>(mz, ma, acc32, (__nv_bfloat16*)dw, sync, tiles_x,
                                                                    tiles_y, n_patches);
    return (int)cudaGetLastError();
}

}  // extern "C"
