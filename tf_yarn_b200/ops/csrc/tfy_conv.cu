// tcgen05 implicit-GEMM 3x3 convolution kernels (C_in = 32 -> C_out = 64, stride 1, VALID) for sm_100a:
// the second convolution of the MNIST-CNN, forward (fused with bias + ReLU + 2x2 max-pool + dropout),
// data gradient (fused with the ReLU gate of the previous layer) and weight gradient.
//
// Shared-memory operand format (all three kernels).  A work item is an 8 x 8 pixel patch of an image
// PAIR, stored as
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
//                                                               (A: MN-major, B: MN-major)
// Weights ([O][3][3][C] bf16) sit in shared memory as [tap][c group][o][8 c]: K-major for fprop and, read
// with the MN-major flag, the transposed operand dgrad needs -- no transpose kernel.
//
// All three are persistent (one CTA per SM walking the patch list) and warp-specialised:
//   producer warps  : cp.async (16 B, L2 -> smem) gathers into a 3-4 stage ring.  The first version of
//                     these kernels used 5-D TMA boxes with a 16-byte inner dimension for the same layout;
//                     ncu showed one L2 request per 16-byte row at ~3 cycles/row/SM (5 B/cycle): 25-33 us
//                     per kernel (profiles/conv_check_tma_v1_r1h.json).  LSU gathers move 512 B per warp
//                     instruction instead; TMA keeps the jobs it is good at (the GEMM's 128-byte rows).
//   MMA warp        : one elected thread issues tcgen05.mma (kind::f16) into a double-buffered TMEM tile
//   epilogue warps  : tcgen05.ld, fused epilogue, overlapped with the next patch's MMAs
#include "tfy_common.cuh"

namespace {

constexpr int CIN = 32, COUT = 64, TAPS = 9;
constexpr int CG_IN = CIN / 8, CG_OUT = COUT / 8;
constexpr int HALO = 10;                               // 8 + 2
constexpr int ROW_B = HALO * 16;                       // 160  : one (h, n) row of a halo patch
constexpr int HROW_B = 2 * ROW_B;                      // 320  : one h step (2 images)
constexpr int HALO_G_B = HALO * HROW_B + 16;           // 3216 : channel-group stride of a halo patch (+16: banks)
constexpr int PATCH_G_B = 8 * 2 * 8 * 16 + 16;         // 2064 : channel-group stride of an 8x8x2 patch
constexpr int W_G_B = COUT * 16 + 16;                  // 1040 : channel-group stride inside a weight tap
constexpr int W_TAP_B = CG_IN * W_G_B;                 // 4160
constexpr int W_BYTES = TAPS * W_TAP_B;                // 37440
constexpr int PRODUCERS = 128;                         // 4 producer warps

constexpr int FP_STAGES = 3, FP_A_B = CG_IN * HALO_G_B;         // 12864
constexpr int FP_EPI_WARPS = 16, FP_THREADS = (FP_EPI_WARPS + 1) * 32 + PRODUCERS;  // 672
constexpr size_t FP_SMEM = (size_t)FP_STAGES * FP_A_B + W_BYTES + 256;
constexpr int DG_STAGES = 3, DG_A_B = CG_OUT * HALO_G_B;        // 25728
constexpr int DG_EPI_WARPS = 4, DG_THREADS = (DG_EPI_WARPS + 1) * 32 + PRODUCERS;   // 288
constexpr size_t DG_SMEM = (size_t)DG_STAGES * DG_A_B + W_BYTES + 256;
constexpr int WG_STAGES = 4, WG_DZ_B = CG_OUT * PATCH_G_B;      // 16512
constexpr int WG_STAGE_B = WG_DZ_B + FP_A_B;                    // 29376
constexpr int WG_THREADS = 288;
constexpr size_t WG_SMEM = (size_t)WG_STAGES * WG_STAGE_B + 256;
constexpr int WG_OUT = COUT * TAPS * CIN;                       // 18432

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void c_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c_smem_u32(bar)) : "memory");
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
// 16-byte asynchronous copy L2 -> shared (L1 bypass); !valid writes 16 zero bytes without touching memory
__device__ __forceinline__ void c_cp16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void c_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void c_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes (cp.async) -> async-proxy reads (tcgen05.mma operand fetch)
__device__ __forceinline__ void c_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
}
__device__ __forceinline__ void c_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
// One lane of a CONVERGED warp.  Unlike `lane == 0`, the compiler knows the guarded region runs on exactly
// one thread and keeps descriptor arithmetic + tcgen05.mma operands in uniform registers; with `lane == 0`
// every UTCHMMA was wrapped in an ELECT / R2UR / BRA.U.ANY loop (~48 cycles per MMA, measured with
// tests/gpu/umma_probe.cu).
__device__ __forceinline__ bool c_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t c_hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// Optional per-CTA event timeline (tests/gpu/conv_check.py --timeline): 16 slots of SM-clock deltas per CTA.
__device__ long long* c_timeline = nullptr;
#define C_MARK(k)                                                                  \
    do {                                                                           \
        if (c_tl) c_tl[(size_t)blockIdx.x * 16 + (k)] = clock64() - c_t0;          \
    } while (0)
#define C_TIMELINE_BEGIN()                 \
    long long* const c_tl = c_timeline;    \
    const long long c_t0 = clock64();      \
    if (c_tl && threadIdx.x == 0) {        \
        unsigned long long ns;             \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns)); \
        c_tl[(size_t)blockIdx.x * 16] = (long long)ns;          \
    }

__device__ __forceinline__ uint8_t* c_align128(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
}

// ---- producers ------------------------------------------------------------------------------------
// Gather plan of one producer thread for a [CG][PH h][2 n][PW w][16 B] patch of an NHWC tensor with CG*8
// channels.  The (source offset, destination offset) pairs depend only on the thread, so they are computed
// once; per patch a slot costs an add, a predicate and the cp.async (the first version recomputed the
// h/n/w split with integer divisions per 16-byte chunk and the producers were the critical path of dgrad).
// Consecutive lanes take consecutive channel groups of a pixel: coalesced 16-byte reads, and the +16 byte
// group stride spreads the writes over the banks.
template <int CG, int PH, int PW, int GSTRIDE, bool BOUNDS>
struct CGather {
    static constexpr int PIX = PH * 2 * PW;
    static constexpr int PER_IT = PRODUCERS / CG;
    static constexpr int NIT = (PIX + PER_IT - 1) / PER_IT;
    int src[NIT];          // element offset from pixel (b0, y0, x0)
    uint32_t dst[NIT];     // byte offset inside the stage; 0xffffffff = no slot
    uint32_t hw[NIT];      // h | w << 8 (bounds checks only)

    __device__ __forceinline__ void init(int ptid, int H, int W) {
        const int g = ptid % CG;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int p = ptid / CG + it * PER_IT;
            const bool valid = p < PIX;
            const int pp = valid ? p : 0;
            const int h = pp / (2 * PW), rem = pp - h * (2 * PW), n = rem / PW, w = rem - n * PW;
            src[it] = ((n * H + h) * W + w) * (CG * 8) + g * 8;
            dst[it] = valid ? (uint32_t)(g * GSTRIDE + (h * 2 + n) * (PW * 16) + w * 16) : 0xffffffffu;
            hw[it] = (uint32_t)h | ((uint32_t)w << 8);
        }
    }
    // tensor: base pointer; (b0, y0, x0): patch origin, possibly outside the image when BOUNDS
    __device__ __forceinline__ void issue(uint32_t stage, const __nv_bfloat16* __restrict__ tensor, int H, int W, int b0,
                                          int y0, int x0) const {
        const long long origin = (((long long)b0 * H + y0) * W + x0) * (CG * 8);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            if (dst[it] != 0xffffffffu) {
                bool ok = true;
                if (BOUNDS) {
                    const int y = y0 + (int)(hw[it] & 0xffu), x = x0 + (int)(hw[it] >> 8);
                    ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
                }
                c_cp16(stage + dst[it], tensor + (ok ? origin + src[it] : 0), ok);
            }
        }
    }
};
using CHaloIn = CGather<CG_IN, HALO, HALO, HALO_G_B, false>;       // activations, always inside the image
using CHaloOut = CGather<CG_OUT, HALO, HALO, HALO_G_B, true>;      // dz with zero padding (dgrad)
using CPatchOut = CGather<CG_OUT, 8, 8, PATCH_G_B, false>;         // dz patch (wgrad)

// weights [O][9][C] -> [tap][c group][o][8 c]
__device__ __forceinline__ void c_load_weights(uint32_t dst, const __nv_bfloat16* __restrict__ w, int ptid) {
#pragma unroll 1
    for (int q = ptid; q < COUT * TAPS * CG_IN; q += PRODUCERS) {
        const int o = q / (TAPS * CG_IN), r = q - o * (TAPS * CG_IN), t = r / CG_IN, g = r - t * CG_IN;
        c_cp16(dst + t * W_TAP_B + g * W_G_B + o * 16, w + (size_t)q * 8, true);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// forward: pooled = dropout(maxpool2x2(relu(conv(a, W) + bias))), code byte per pooled element
// (bits 0-1 = argmax position dy*2+dx inside the 2x2 window, bit 2 = gradient flows).
// warps 0..15 epilogue (TMEM quadrant = warp & 3, 16-column slice = warp >> 2), warp 16 MMA, warps 17..20 producers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FP_THREADS, 1)
tfy_conv3x3_fprop_pool_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ wgt,
                              const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ pooled,
                              uint8_t* __restrict__ code, int H, int W, int tiles_x, int tiles_y, int n_patches,
                              float drop_rate, uint32_t seed, const TfyOptHyper* __restrict__ hp) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align128(smem_raw);
    uint8_t* w_tile = stages + (size_t)FP_STAGES * FP_A_B;
    uint64_t* full = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);
    uint64_t* empty = full + FP_STAGES;
    uint64_t* tfull = empty + FP_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FP_STAGES; ++s) { c_mbar_init(&full[s], PRODUCERS); c_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { c_mbar_init(&tfull[b], 1); c_mbar_init(&tempty[b], FP_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FP_EPI_WARPS) c_tmem_alloc<128>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) C_MARK(1);
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int tiles = tiles_x * tiles_y;

    if (warp > FP_EPI_WARPS) {
        // ---------------- producers
        const int ptid = threadIdx.x - (FP_EPI_WARPS + 1) * 32;
        constexpr int LAG = FP_STAGES - 1;
        CHaloIn plan;
        plan.init(ptid, H, W);
        for (int i = 0; i < my_patches; ++i) {
            const int s = i % FP_STAGES;
            if (i >= FP_STAGES) c_mbar_wait(&empty[s], ((i / FP_STAGES) - 1) & 1);
            const int p = (int)blockIdx.x + i * (int)gridDim.x;
            const int bz = p / tiles, r = p - bz * tiles, ty = r / tiles_x, tx = r - ty * tiles_x;
            if (i == 0) c_load_weights(c_smem_u32(w_tile), wgt, ptid);
            plan.issue(c_smem_u32(stages + (size_t)s * FP_A_B), a, H, W, bz * 2, ty * 8, tx * 8);
            c_cp_commit();
            if (i == 0 && ptid == 0) C_MARK(2);
            if (i >= LAG) {
                c_cp_wait<LAG>();
                c_fence_async_smem();
                if (i == LAG && ptid == 0) C_MARK(3);
                c_mbar_arrive(&full[(i - LAG) % FP_STAGES]);
            }
        }
        c_cp_wait<0>();
        c_fence_async_smem();
        if (ptid == 0) C_MARK(12);
        for (int i = (my_patches > LAG ? my_patches - LAG : 0); i < my_patches; ++i) c_mbar_arrive(&full[i % FP_STAGES]);
    } else if (warp == FP_EPI_WARPS) {
        // ---------------- MMA issuer
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(128, COUT, 0, 0);
            const uint32_t w0 = c_smem_u32(w_tile);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % FP_STAGES, b = i & 1;
                c_mbar_wait(&full[s], (i / FP_STAGES) & 1);
                if (i == 0) C_MARK(4);
                c_fence_async_smem();
                if (i >= 2) c_mbar_wait(&tempty[b], ((i >> 1) - 1) & 1);
                c_fence_after();
                const uint32_t a0 = c_smem_u32(stages + (size_t)s * FP_A_B);
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks) {
                        // A: rows = pixels (8 w per core matrix, (h,n) groups ROW_B apart), K halves = channel groups
                        const uint64_t adesc = c_desc(a0 + kh * HROW_B + kw * 16 + ks * 2 * HALO_G_B, HALO_G_B, ROW_B);
                        // B: rows = output channels (16 B apart, groups of 8 = 128 B), K halves = channel groups
                        const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 2 * W_G_B, W_G_B, 128);
                        c_umma(tmem_base + (uint32_t)(b * COUT), adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                c_commit(&empty[s]);
                c_commit(&tfull[b]);
                if (i == 0) C_MARK(5);
                if (i == my_patches - 1) C_MARK(11);
            }
        }
    } else {
        // ---------------- epilogue: accumulator row r = (h, n, w) = (r/16, (r/8)&1, r%8) lives in TMEM lane r
        const int quad = warp & 3, cq = warp >> 2;
        const int r = quad * 32 + lane;
        const int h = r >> 4, n = (r >> 3) & 1, w = r & 7;
        const int dx = lane & 1, dy = (lane >> 4) & 1;            // position inside the 2x2 pooling window
        const uint32_t tag = (uint32_t)(dy * 2 + dx);
        const int OH = H - 2, OW = W - 2, PH = OH / 2, PW = OW / 2;
        const uint32_t step = hp ? (uint32_t)hp->step : 0u;
        // dropout decisions come 4 to a hash (one byte each): the rate is quantised to 1/256
        const uint32_t thr = drop_rate > 0.f ? (uint32_t)(drop_rate * 256.f + 0.5f) : 0u;
        const float keep_scale = 256.f / (256.f - (float)thr);
        const uint32_t salt = seed ^ c_hash32(step * 0x9E3779B9U + 0x85ebca6bU);
        // after the two exchanges below lane (dy, dx) owns columns [8 dx + 4 dy, +4) of the warp's 16
        const int col = cq * 16 + dx * 8 + dy * 4;
        const uint2 braw = *reinterpret_cast<const uint2*>(bias + col);
        const float bv[4] = {tfy_bf16lo(braw.x), tfy_bf16hi(braw.x), tfy_bf16lo(braw.y), tfy_bf16hi(braw.y)};
        for (int i = 0; i < my_patches; ++i) {
            const int b = i & 1;
            const int p = (int)blockIdx.x + i * (int)gridDim.x;
            const int bz = p / tiles, rr = p - bz * tiles, ty = rr / tiles_x, tx = rr - ty * tiles_x;
            const size_t prow = (((size_t)(bz * 2 + n) * PH + (ty * 8 + h) / 2) * PW + (tx * 8 + w) / 2) * COUT;
            c_mbar_wait(&tfull[b], (i >> 1) & 1);
            if (threadIdx.x == 0) { if (i == 0) C_MARK(6); if (i == my_patches - 1) C_MARK(8); }
            c_fence_after();
            uint32_t acc[16];
            c_tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * COUT + cq * 16), acc);
            c_tmem_ld_wait();
            c_fence_before();
            c_mbar_arrive(&tempty[b]);                             // the MMA warp may overwrite this buffer now
            // tag every value with its window position in the 2 low mantissa bits, then reduce-scatter the
            // max over the window (lane ^ 1: other column, lane ^ 16: other row)
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float((acc[j] & ~3u) | tag);
            float m8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float send = dx ? v[j] : v[8 + j];
                const float keep = dx ? v[8 + j] : v[j];
                m8[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
            }
            float m4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float send = dy ? m8[j] : m8[4 + j];
                const float keep = dy ? m8[4 + j] : m8[j];
                m4[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            }
            const uint32_t rnd =
                thr ? c_hash32(salt ^ c_hash32((uint32_t)((prow + col) >> 2) * 0xC2B2AE35U + 0x27d4eb2fU)) : 0xffffffffu;
            float o[4];
            uint32_t codes = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t pos = __float_as_uint(m4[j]) & 3u;
                float x = m4[j] + bv[j];
                const bool on = x > 0.f && ((rnd >> (8 * j)) & 0xffu) >= thr;
                x = on ? x * keep_scale : 0.f;
                o[j] = x;
                codes |= (pos | (on ? 4u : 0u)) << (8 * j);
            }
            __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
            uint2 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&lo);
            packed.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(pooled + prow + col) = packed;
            *reinterpret_cast<uint32_t*>(code + prow + col) = codes;
            if (threadIdx.x == 0) { if (i == 0) C_MARK(7); if (i == my_patches - 1) C_MARK(9); }
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == FP_EPI_WARPS) c_tmem_free<128>(tmem_base);
    if (threadIdx.x == 0) C_MARK(10);
}

// ------------------------------------------------------------------------------------------------
// data gradient: dx[b, y, x, c] = sum_{kh,kw,o} dz[b, y-kh, x-kw, o] W[o, kh, kw, c], optionally gated by
// the previous layer's ReLU (gate = that layer's output: dx is zeroed where gate <= 0).
// dz: [B, H-2, W-2, 64]; dx: [B, H, W, 32].  warps 0..3 epilogue, warp 4 MMA, warps 5..8 producers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DG_THREADS, 1)
tfy_conv3x3_dgrad_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ wgt,
                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ dx, int H, int W,
                         int tiles_x, int tiles_y, int n_patches) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align128(smem_raw);
    uint8_t* w_tile = stages + (size_t)DG_STAGES * DG_A_B;
    uint64_t* full = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);
    uint64_t* empty = full + DG_STAGES;
    uint64_t* tfull = empty + DG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DG_STAGES; ++s) { c_mbar_init(&full[s], PRODUCERS); c_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { c_mbar_init(&tfull[b], 1); c_mbar_init(&tempty[b], DG_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == DG_EPI_WARPS) c_tmem_alloc<64>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) C_MARK(1);
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int tiles = tiles_x * tiles_y;

    if (warp > DG_EPI_WARPS) {
        const int ptid = threadIdx.x - (DG_EPI_WARPS + 1) * 32;
        constexpr int LAG = DG_STAGES - 1;
        CHaloOut plan;
        plan.init(ptid, H - 2, W - 2);
        for (int i = 0; i < my_patches; ++i) {
            const int s = i % DG_STAGES;
            if (i >= DG_STAGES) c_mbar_wait(&empty[s], ((i / DG_STAGES) - 1) & 1);
            const int p = (int)blockIdx.x + i * (int)gridDim.x;
            const int bz = p / tiles, r = p - bz * tiles, ty = r / tiles_x, tx = r - ty * tiles_x;
            if (i == 0) c_load_weights(c_smem_u32(w_tile), wgt, ptid);
            // halo origin (y0-2, x0-2) of the dz image (H-2 x W-2): pixels outside it read as zero
            plan.issue(c_smem_u32(stages + (size_t)s * DG_A_B), dz, H - 2, W - 2, bz * 2, ty * 8 - 2, tx * 8 - 2);
            c_cp_commit();
            if (i == 0 && ptid == 0) C_MARK(2);
            if (i >= LAG) {
                c_cp_wait<LAG>();
                c_fence_async_smem();
                if (i == LAG && ptid == 0) C_MARK(3);
                c_mbar_arrive(&full[(i - LAG) % DG_STAGES]);
            }
        }
        c_cp_wait<0>();
        c_fence_async_smem();
        if (ptid == 0) C_MARK(12);
        for (int i = (my_patches > LAG ? my_patches - LAG : 0); i < my_patches; ++i) c_mbar_arrive(&full[i % DG_STAGES]);
    } else if (warp == DG_EPI_WARPS) {
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(128, CIN, 0, 1);                    // B is MN-major: N = c contiguous
            const uint32_t w0 = c_smem_u32(w_tile);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % DG_STAGES, b = i & 1;
                c_mbar_wait(&full[s], (i / DG_STAGES) & 1);
                if (i == 0) C_MARK(4);
                c_fence_async_smem();
                if (i >= 2) c_mbar_wait(&tempty[b], ((i >> 1) - 1) & 1);
                c_fence_after();
                const uint32_t a0 = c_smem_u32(stages + (size_t)s * DG_A_B);
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
#pragma unroll
                    for (int ks = 0; ks < COUT / 16; ++ks) {
                        // dz[y-kh, x-kw] sits (2-kh, 2-kw) into the halo patch
                        const uint64_t adesc =
                            c_desc(a0 + (2 - kh) * HROW_B + (2 - kw) * 16 + ks * 2 * HALO_G_B, HALO_G_B, ROW_B);
                        // W tap as [c group][o][8 c]: K = o (16 B apart, groups of 8 o = 128 B), N groups = c groups
                        const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 16 * 16, 128, W_G_B);
                        c_umma(tmem_base + (uint32_t)(b * CIN), adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                c_commit(&empty[s]);
                c_commit(&tfull[b]);
                if (i == 0) C_MARK(5);
                if (i == my_patches - 1) C_MARK(11);
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const int h = r >> 4, n = (r >> 3) & 1, w = r & 7;
        for (int i = 0; i < my_patches; ++i) {
            const int b = i & 1;
            const int p = (int)blockIdx.x + i * (int)gridDim.x;
            const int bz = p / tiles, rr = p - bz * tiles, ty = rr / tiles_x, tx = rr - ty * tiles_x;
            c_mbar_wait(&tfull[b], (i >> 1) & 1);
            if (threadIdx.x == 0) { if (i == 0) C_MARK(6); if (i == my_patches - 1) C_MARK(8); }
            c_fence_after();
            uint32_t acc[2][16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * CIN);
            c_tmem_ld16(taddr, acc[0]);
            c_tmem_ld16(taddr + 16, acc[1]);
            c_tmem_ld_wait();
            c_fence_before();
            c_mbar_arrive(&tempty[b]);
            const int y = ty * 8 + h, x = tx * 8 + w;
            if (y < H && x < W) {
                const size_t off = (((size_t)(bz * 2 + n) * H + y) * W + x) * CIN;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[ch][j]);
                    if (gate) {
                        float g[16];
                        TfyPack<__nv_bfloat16>::unpack(tfy_ld16(gate + off + ch * 16), g);
                        TfyPack<__nv_bfloat16>::unpack(tfy_ld16(gate + off + ch * 16 + 8), g + 8);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = g[j] > 0.f ? v[j] : 0.f;
                    }
                    tfy_st16(dx + off + ch * 16, TfyPack<__nv_bfloat16>::pack(v));
                    tfy_st16(dx + off + ch * 16 + 8, TfyPack<__nv_bfloat16>::pack(v + 8));
                }
            }
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == DG_EPI_WARPS) c_tmem_free<64>(tmem_base);
    if (threadIdx.x == 0) C_MARK(10);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[o, t, c] = sum_{b,y,x} dz[b, y, x, o] a[b, y+kh, x+kw, c].  The pixel patches are the
// GEMM K dimension: every CTA accumulates nine 64x32 tiles in TMEM (one per tap: 288 of the 512 columns)
// over its share of the patches, stores the partial to `partials[blockIdx.x]` (fp32, L2 resident), and after
// a grid-wide arrive/release reduces a slice of the 18432 outputs over all partials in a fixed order
// (deterministic, no atomics) and writes bf16 dW.
// warps 0..3 epilogue, warp 4 MMA, warps 5..8 producers.  sync: two uint32, zero-initialised once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WG_THREADS, 1)
tfy_conv3x3_wgrad_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ dz,
                         float* __restrict__ partials, __nv_bfloat16* __restrict__ dw, uint32_t* __restrict__ sync,
                         int H, int W, int tiles_x, int tiles_y, int n_patches) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align128(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)WG_STAGES * WG_STAGE_B);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tmem_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gen0 = threadIdx.x == 0 ? *reinterpret_cast<volatile uint32_t*>(sync + 1) : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { c_mbar_init(&full[s], PRODUCERS); c_mbar_init(&empty[s], 1); }
        c_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) c_tmem_alloc<512>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) C_MARK(1);
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int tiles = tiles_x * tiles_y;

    if (warp > 4) {
        const int ptid = threadIdx.x - 5 * 32;
        constexpr int LAG = WG_STAGES - 1;
        CPatchOut plan_dz;
        CHaloIn plan_a;
        plan_dz.init(ptid, H - 2, W - 2);
        plan_a.init(ptid, H, W);
        for (int i = 0; i < my_patches; ++i) {
            const int s = i % WG_STAGES;
            if (i >= WG_STAGES) c_mbar_wait(&empty[s], ((i / WG_STAGES) - 1) & 1);
            const int p = (int)blockIdx.x + i * (int)gridDim.x;
            const int bz = p / tiles, r = p - bz * tiles, ty = r / tiles_x, tx = r - ty * tiles_x;
            const uint32_t st = c_smem_u32(stages + (size_t)s * WG_STAGE_B);
            plan_dz.issue(st, dz, H - 2, W - 2, bz * 2, ty * 8, tx * 8);
            plan_a.issue(st + WG_DZ_B, a, H, W, bz * 2, ty * 8, tx * 8);
            c_cp_commit();
            if (i == 0 && ptid == 0) C_MARK(2);
            if (i >= LAG) {
                c_cp_wait<LAG>();
                c_fence_async_smem();
                if (i == LAG && ptid == 0) C_MARK(3);
                c_mbar_arrive(&full[(i - LAG) % WG_STAGES]);
            }
        }
        c_cp_wait<0>();
        c_fence_async_smem();
        if (ptid == 0) C_MARK(12);
        for (int i = (my_patches > LAG ? my_patches - LAG : 0); i < my_patches; ++i) c_mbar_arrive(&full[i % WG_STAGES]);
    } else if (warp == 4) {
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(64, CIN, 1, 1);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % WG_STAGES;
                c_mbar_wait(&full[s], (i / WG_STAGES) & 1);
                if (i == 0) C_MARK(4);
                c_fence_async_smem();
                c_fence_after();
                const uint32_t dz0 = c_smem_u32(stages + (size_t)s * WG_STAGE_B), a0 = dz0 + WG_DZ_B;
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
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
            C_MARK(5);
        }
    } else if (my_patches > 0) {
        // M = 64 accumulator: row o lives in TMEM lane (o % 16) + 32 * (o / 16): lanes 0..15 of each quadrant
        const int quad = warp & 3;
        c_mbar_wait(tmem_full, 0);
        if (threadIdx.x == 0) C_MARK(6);
        c_fence_after();
        float* mine = partials + (size_t)blockIdx.x * WG_OUT + (size_t)(quad * 16 + (lane & 15)) * (TAPS * CIN);
#pragma unroll 1
        for (int col = 0; col < TAPS * CIN; col += 32) {
            uint32_t acc[2][16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col;
            c_tmem_ld16(taddr, acc[0]);
            c_tmem_ld16(taddr + 16, acc[1]);
            c_tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        __stcg(reinterpret_cast<float4*>(mine + col + ch * 16 + j),
                               make_float4(__uint_as_float(acc[ch][j]), __uint_as_float(acc[ch][j + 1]),
                                           __uint_as_float(acc[ch][j + 2]), __uint_as_float(acc[ch][j + 3])));
            }
        }
        __threadfence();
    }
    c_fence_before();
    __syncthreads();
    if (warp == 4) c_tmem_free<512>(tmem_base);
    if (threadIdx.x == 0) C_MARK(7);

    // grid-wide arrive / release (all CTAs are co-resident: grid <= #SMs, one CTA per SM)
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t arrived = atomicAdd(sync, 1u);
        if (arrived == gridDim.x - 1) {
            *reinterpret_cast<volatile uint32_t*>(sync) = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        } else {
            while (*reinterpret_cast<volatile uint32_t*>(sync + 1) == gen0) __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) C_MARK(8);
    // slice of the outputs owned by this CTA (float4 units), summed over the partials in a fixed order
    constexpr int TOTAL4 = WG_OUT / 4;
    const int per = (TOTAL4 + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * per, hi = min(lo + per, TOTAL4);
    const int len = max(hi - lo, 0);
    const int n_part = (int)gridDim.x;                               // every CTA wrote a partial
    const float4* src = reinterpret_cast<const float4*>(partials) + lo;
    auto store = [&](int f, const float4& s4) {
        __nv_bfloat162 l2 = __floats2bfloat162_rn(s4.x, s4.y), h2 = __floats2bfloat162_rn(s4.z, s4.w);
        uint2 packed;
        packed.x = *reinterpret_cast<uint32_t*>(&l2);
        packed.y = *reinterpret_cast<uint32_t*>(&h2);
        *reinterpret_cast<uint2*>(dw + (size_t)(lo + f) * 4) = packed;
    };
    if (len > 0 && 2 * len <= WG_THREADS) {
        // few outputs per CTA (the usual case: 32): `parts` thread groups split the list of partials
        float4* red = reinterpret_cast<float4*>(stages);             // the ring is idle now
        const int parts = min(WG_THREADS / len, 16);
        const int f = (int)threadIdx.x % len, part = (int)threadIdx.x / len;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (part < parts) {
#pragma unroll 8
            for (int q = part; q < n_part; q += parts) {
                const float4 v = __ldcg(src + (size_t)q * TOTAL4 + f);
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
        }
        red[threadIdx.x] = s4;
        __syncthreads();
        if (part == 0) {
            for (int q = 1; q < parts; ++q) {
                const float4 v = red[q * len + f];
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
            store(f, s4);
        }
    } else {
        for (int f = (int)threadIdx.x; f < len; f += WG_THREADS) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int q = 0; q < n_part; ++q) {
                const float4 v = __ldcg(src + (size_t)q * TOTAL4 + f);
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
            store(f, s4);
        }
    }
    if (threadIdx.x == 0) C_MARK(10);
}

namespace {
bool c_attr_set = false;
int c_sms = 0;
bool c_init() {
    if (!c_attr_set) {
        if (cudaFuncSetAttribute(tfy_conv3x3_fprop_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)FP_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)DG_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)WG_SMEM) != cudaSuccess)
            return false;
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&c_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c_sms <= 0)
            return false;
        c_attr_set = true;
    }
    return true;
}
// persistent grid: one CTA per SM, trimmed so that the patches divide as evenly as possible
int c_grid(int n_patches) {
    if (n_patches <= c_sms) return n_patches;
    const int waves = (n_patches + c_sms - 1) / c_sms;
    return (n_patches + waves - 1) / waves;
}
}  // namespace

extern "C" {

// debug: device buffer of 16 x int64 per CTA receiving the event timeline of the next launches (nullptr = off)
int tfy_conv_set_timeline(long long* buf) {
    return (int)cudaMemcpyToSymbol(c_timeline, &buf, sizeof(buf));
}

// number of fp32 elements of the `partials` scratch buffer tfy_conv3x3_c32_wgrad needs
size_t tfy_conv3x3_c32_wgrad_scratch_elems() { return (size_t)160 * WG_OUT; }

// a: [B, H, W, 32] bf16 (NHWC), w: [64, 3, 3, 32] bf16, bias: [64] bf16
// pooled / code: [B, (H-2)/2, (W-2)/2, 64].  Requires (H-2) % 8 == 0, (W-2) % 8 == 0, B % 2 == 0.
int tfy_conv3x3_c32_pool_fwd(const void* a, const void* w, const void* bias, void* pooled, void* code, int B, int H,
                             int W, float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2) || drop_rate < 0.f || drop_rate >= 1.f) return -2;
    if (!c_init()) return -4;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_conv3x3_fprop_pool_kernel<<<c_grid(n_patches), FP_THREADS, FP_SMEM, s>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)pooled,
        (uint8_t*)code, H, W, tiles_x, tiles_y, n_patches, drop_rate, seed, hp);
    return (int)cudaGetLastError();
}

// dz: [B, H-2, W-2, 64] bf16, w: [64, 3, 3, 32], gate (optional): [B, H, W, 32], dx: [B, H, W, 32].  B % 2 == 0.
int tfy_conv3x3_c32_dgrad(const void* dz, const void* w, const void* gate, void* dx, int B, int H, int W,
                          cudaStream_t s) {
    if (B % 2) return -2;
    if (!c_init()) return -4;
    const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_conv3x3_dgrad_kernel<<<c_grid(n_patches), DG_THREADS, DG_SMEM, s>>>(
        (const __nv_bfloat16*)dz, (const __nv_bfloat16*)w, (const __nv_bfloat16*)gate, (__nv_bfloat16*)dx, H, W,
        tiles_x, tiles_y, n_patches);
    return (int)cudaGetLastError();
}

// a: [B, H, W, 32], dz: [B, H-2, W-2, 64], dw: [64, 3, 3, 32] bf16 (overwritten).
// partials: tfy_conv3x3_c32_wgrad_scratch_elems() floats of scratch; sync: 2 x uint32, zeroed once at allocation.
int tfy_conv3x3_c32_wgrad(const void* a, const void* dz, float* partials, void* dw, uint32_t* sync, int B, int H, int W,
                          cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2)) return -2;
    if (!c_init() || c_sms > 160) return -4;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_conv3x3_wgrad_kernel<<<c_grid(n_patches), WG_THREADS, WG_SMEM, s>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)dz, partials, (__nv_bfloat16*)dw, sync, H, W, tiles_x, tiles_y,
        n_patches);
    return (int)cudaGetLastError();
}

}  // extern "C"
