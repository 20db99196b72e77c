// tcgen05 implicit-GEMM 3x3 convolution kernels (C_in = 32 -> C_out = 64, stride 1, VALID) for sm_100a:
// the second convolution of the MNIST-CNN, forward (fused with bias + ReLU + 2x2 max-pool + dropout),
// data gradient (fused with the ReLU gate of the previous layer) and weight gradient.
//
// Shared-memory operand format (all three kernels).  A work item is an 8 x 8 pixel patch of an image
// PAIR; TMA stores the (haloed) patch as
//        [h][image n][w][all channels of the pixel]         one pixel = one 64- or 128-byte row,
// with the 64B / 128B hardware swizzle.  tests/gpu/umma_probe.cu established that the UMMA operand fetch
// XORs the 16-byte chunk index with ABSOLUTE shared-memory address bits (base_offset = 0), whatever the
// descriptor's start address and stride between 8-row groups are.  So a filter tap (kh, kw) is just a
// different START ADDRESS into the same halo patch (+kh rows, +kw pixels) and the stride between 8-pixel
// groups is the 10-pixel halo row: the halo is loaded ONCE per patch by ONE TMA and the nine im2col views
// never materialise, neither in HBM nor in shared memory.  The pixel dimension is GEMM-M for fprop / dgrad
// (K-major A operand) and GEMM-K for wgrad (MN-major operands: the SAME bytes, other major-ness).
//
//   fprop : D[pix, o]  = sum_{t,c} a[pix + t, c] W[o, t, c]     M=128 N=64 K=9x32   (A: K-major, B: K-major)
//   dgrad : D[pix, c]  = sum_{t,o} dz[pix - t, o] W[o, t, c]    M=128 N=32 K=9x64   (A: K-major, B: MN-major)
//   wgrad : D[o, (t,c)] = sum_pix dz[pix, o] a[pix + t, c]      M=64  N=32 K=128/patch, 9 accumulators
//                                                               (A: MN-major, B: MN-major)
// Weights ([O][3][3][C] bf16) are loaded by one TMA as [tap][o][32 c] (64-byte rows): K-major for fprop
// and, read with the MN-major flag, the transposed operand dgrad needs -- no transpose kernel.
//
// All three are persistent (one CTA per SM walking the patch list) and warp-specialised: one TMA thread,
// one MMA thread (elect.sync, so descriptors stay in uniform registers and UTCHMMAs issue back to back),
// epilogue warps overlapped with the next patch through a double-buffered TMEM accumulator.
//
// History, kept because the numbers drove the design (profiles/conv_check_*.json, umma_probe_*.jsonl):
//   v1  5-D TMA boxes with a 16-byte inner dimension into an un-swizzled layout: one L2 request per
//       16-byte row at ~3 cycles/row/SM -> 25-33 us per kernel.
//   v2  cp.async gathers for the same layout: LDGSTS.128 sustains ~14 B/cycle/SM and the first operands
//       land 5.6 us after launch -> 13-20 us per kernel.
//   v3  (this file) whole-pixel rows (64/128 B) + swizzle: 200 TMA rows per patch instead of 800-1600.
#include <cuda.h>

#include "tfy_common.cuh"
#include "tfy_fused_step.cuh"

namespace {

constexpr int CIN = 32, COUT = 64, TAPS = 9;
constexpr int HALO = 10;                               // 8 + 2
constexpr int A_PIX_B = CIN * 2;                       // 64  : one input pixel  (SWIZZLE_64B rows)
constexpr int Z_PIX_B = COUT * 2;                      // 128 : one output pixel (SWIZZLE_128B rows)
constexpr int A_ROW_B = HALO * A_PIX_B;                // 640  : one (h, n) row of an input halo patch
constexpr int Z_ROW_B = HALO * Z_PIX_B;                // 1280 : one (h, n) row of an output halo patch
constexpr int A_HALO_B = HALO * 2 * A_ROW_B;           // 12800
constexpr int A_HALO_PAD = 13312;                      // stage stride: keeps every TMA destination 1024-aligned
constexpr int Z_HALO_B = HALO * 2 * Z_ROW_B;           // 25600 (= 25 x 1024)
constexpr int Z_PATCH_B = 8 * 2 * 8 * Z_PIX_B;         // 16384
constexpr int W_TAP_B = COUT * A_PIX_B;                // 4096 : [64 o][32 c]
constexpr int W_BYTES = TAPS * W_TAP_B;                // 36864

constexpr int FP_STAGES = 3, FP_EPI_WARPS = 16, FP_THREADS = (FP_EPI_WARPS + 2) * 32;     // 576
constexpr size_t FP_SMEM = (size_t)FP_STAGES * A_HALO_PAD + W_BYTES + 1024 + 256;
constexpr int DG_STAGES = 3, DG_EPI_WARPS = 4, DG_THREADS = (DG_EPI_WARPS + 2 + 4) * 32;  // 320: +4 un-pool producer warps
constexpr size_t DG_SMEM = (size_t)DG_STAGES * Z_HALO_B + W_BYTES + 1024 + 256;
constexpr int WG_STAGES = 4, WG_STAGE_B = Z_PATCH_B + A_HALO_PAD;                         // 29696
constexpr int WG_WORK_WARPS = 16, WG_THREADS = (WG_WORK_WARPS + 2) * 32;                   // 576
constexpr int WG_OUT = COUT * TAPS * CIN;                                                 // 18432
constexpr int WG_PART = WG_OUT + COUT;                 // per-CTA partial: dW tile + db (un-pool variant)
constexpr size_t WG_SMEM = (size_t)WG_STAGES * WG_STAGE_B + 1024 + 512 + WG_WORK_WARPS * COUT * 4;

constexpr uint32_t SW128 = 2, SW64 = 4;                // UMMA descriptor layout types

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void c_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c_smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void c_tma_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                         int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(c_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void c_tma_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(c_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void c_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Swizzled UMMA shared-memory descriptor (base_offset 0: the swizzle phase comes from the absolute address).
// sbo = byte distance between consecutive groups of 8 rows (K-major: 8 M/N rows; MN-major: 8 K rows).
__device__ __forceinline__ uint64_t c_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)layout << 61;
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
// Optional per-CTA event timeline (tests/gpu/conv_check.py): 16 slots of SM-clock deltas per CTA.
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

__device__ __forceinline__ uint8_t* c_align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

#include "tfy_conv_index.cuh"

// Un-pooling producer task: the gradient dp of ONE pooled pixel (8 channels of group g) becomes the 2x2 window
// of dz rows it came from, written straight into a swizzled [pixel rows][128 B] operand tile:
//   dz[2py+dy, 2px+dx, o] = (code[o] & 4) && (code[o] & 3) == dy*2+dx ? dp[o] * scale : 0
// (code byte of the forward kernel: argmax position + gate).  The fused pool/dropout/ReLU backward never
// writes the 4x larger dz tensor to memory.  row0 = tile row of the window's top-left pixel, row_dy = rows per
// image row of the tile (2 images interleaved: 2 * pixels per row).  Returns the gated values for db.
struct CUnpoolIn {
    uint4 v;
    uint2 cd;
};
__device__ __forceinline__ CUnpoolIn c_unpool_load(const __nv_bfloat16* __restrict__ dp,
                                                   const uint8_t* __restrict__ code, bool valid, size_t pidx, int g) {
    CUnpoolIn in;
    in.v = make_uint4(0, 0, 0, 0);
    in.cd = make_uint2(0, 0);
    if (valid) {
        in.v = tfy_ld16(dp + pidx * COUT + g * 8);
        in.cd = *reinterpret_cast<const uint2*>(code + pidx * COUT + g * 8);
    }
    return in;
}
__device__ __forceinline__ void c_unpool_emit(uint8_t* tile, const CUnpoolIn& in, float scale, int g, int row0,
                                              int row_dy, float* gated) {
    float f[8];
    TfyPack<__nv_bfloat16>::unpack(in.v, f);
    uint32_t c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        c[k] = ((k < 4 ? in.cd.x >> (8 * k) : in.cd.y >> (8 * (k - 4)))) & 0xffu;
        f[k] = (c[k] & 4u) ? f[k] * scale : 0.f;
        gated[k] = f[k];
    }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ((c[k] & 3u) == (uint32_t)pos) ? f[k] : 0.f;
        const int row = row0 + (pos >> 1) * row_dy + (pos & 1);
        *reinterpret_cast<uint4*>(tile + (size_t)row * Z_PIX_B + ((uint32_t)(g ^ (row & 7)) << 4)) =
            TfyPack<__nv_bfloat16>::pack(o);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// forward: pooled = dropout(maxpool2x2(relu(conv(a, W) + bias))), code byte per pooled element
// (bits 0-1 = argmax position dy*2+dx inside the 2x2 window, bit 2 = gradient flows).
// warps 0..15 epilogue (TMEM quadrant = warp & 3, 16-column slice = warp >> 2), warp 16 MMA, warp 17 TMA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FP_THREADS, 1)
tfy_conv3x3_fprop_pool_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                              const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ pooled,
                              uint8_t* __restrict__ code, int H, int W, int tiles_x, int tiles_y, int n_patches,
                              float drop_rate, uint32_t seed, const TfyOptHyper* __restrict__ hp) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align1024(smem_raw);                      // [3][10 h][2 n][10 w][64 B], 64B swizzle
    uint8_t* w_tile = stages + (size_t)FP_STAGES * A_HALO_PAD;    // [9][64 o][64 B], 64B swizzle
    uint64_t* full = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);
    uint64_t* empty = full + FP_STAGES;
    uint64_t* tfull = empty + FP_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        c_prefetch_map(&map_a);
        c_prefetch_map(&map_w);
        for (int s = 0; s < FP_STAGES; ++s) { c_mbar_init(&full[s], 1); c_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { c_mbar_init(&tfull[b], 1); c_mbar_init(&tempty[b], FP_EPI_WARPS * 32); }
        c_mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FP_EPI_WARPS) c_tmem_alloc<128>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();                      // upstream grids complete; our dependents may start launching
    if (threadIdx.x == 0) C_MARK(1);
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == FP_EPI_WARPS + 1) {
        // ---------------- TMA producer
        if (c_elect_one()) {
            c_mbar_expect_tx(wfull, W_BYTES);
            c_tma_3d(&map_w, wfull, w_tile, 0, 0, 0);
            CPatchIter it(tiles_x, tiles_y);
            for (int i = 0; i < my_patches; ++i, it.next()) {
                const int s = i % FP_STAGES;
                if (i >= FP_STAGES) c_mbar_wait(&empty[s], ((i / FP_STAGES) - 1) & 1);
                c_mbar_expect_tx(&full[s], A_HALO_B);
                c_tma_4d(&map_a, &full[s], stages + (size_t)s * A_HALO_PAD, 0, it.tx * 8, it.bz * 2, it.ty * 8);
                if (i == 0) C_MARK(2);
            }
        }
    } else if (warp == FP_EPI_WARPS) {
        // ---------------- MMA issuer
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(128, COUT, 0, 0);
            const uint32_t w0 = c_smem_u32(w_tile);
            c_mbar_wait(wfull, 0);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % FP_STAGES, b = i & 1;
                c_mbar_wait(&full[s], (i / FP_STAGES) & 1);
                if (i == 0) C_MARK(4);
                if (i >= 2) c_mbar_wait(&tempty[b], ((i >> 1) - 1) & 1);
                c_fence_after();
                const uint32_t a0 = c_smem_u32(stages + (size_t)s * A_HALO_PAD);
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks) {
                        // A rows = pixels: 8 consecutive w (64 B apart), next group = next (h, n) row of the halo
                        const uint64_t adesc = c_desc(a0 + kh * 2 * A_ROW_B + kw * A_PIX_B + ks * 32, A_ROW_B, SW64);
                        // B rows = output channels (64 B apart), 8-row groups 512 B apart
                        const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 32, 8 * A_PIX_B, SW64);
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
        CPatchIter it(tiles_x, tiles_y);
        for (int i = 0; i < my_patches; ++i, it.next()) {
            const int b = i & 1;
            const uint32_t prow =
                (uint32_t)((((it.bz * 2 + n) * PH + ((it.ty * 8 + h) >> 1)) * PW + ((it.tx * 8 + w) >> 1)) * COUT + col);
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
            const uint32_t rnd = thr ? c_hash32(salt ^ c_hash32((prow >> 2) * 0xC2B2AE35U + 0x27d4eb2fU)) : 0xffffffffu;
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
            *reinterpret_cast<uint2*>(pooled + prow) = packed;
            *reinterpret_cast<uint32_t*>(code + prow) = codes;
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
// dz: [B, H-2, W-2, 64]; dx: [B, H, W, 32].  warps 0..3 epilogue, warp 4 MMA, warp 5 TMA (zero-fills the halo)
// ------------------------------------------------------------------------------------------------
// UNPOOL: the A operand is not loaded from a materialised dz but rebuilt by warps 6..9 from the pooled gradient
// dp [B, (H-2)/2, (W-2)/2, 64] and the forward kernel's code bytes (c_unpool_task).
template <bool UNPOOL>
__global__ void __launch_bounds__(DG_THREADS, 1)
tfy_conv3x3_dgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_w,
                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ dx, int H, int W,
                         int tiles_x, int tiles_y, int n_patches, const __nv_bfloat16* __restrict__ dp,
                         const uint8_t* __restrict__ code, float scale) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align1024(smem_raw);                      // [3][10 h][2 n][10 w][128 B], 128B swizzle
    uint8_t* w_tile = stages + (size_t)DG_STAGES * Z_HALO_B;      // [9][64 o][64 B], 64B swizzle
    uint64_t* full = reinterpret_cast<uint64_t*>(w_tile + W_BYTES);
    uint64_t* empty = full + DG_STAGES;
    uint64_t* tfull = empty + DG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        if (!UNPOOL) c_prefetch_map(&map_dz);
        c_prefetch_map(&map_w);
        for (int s = 0; s < DG_STAGES; ++s) { c_mbar_init(&full[s], UNPOOL ? 128 : 1); c_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { c_mbar_init(&tfull[b], 1); c_mbar_init(&tempty[b], DG_EPI_WARPS * 32); }
        c_mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == DG_EPI_WARPS) c_tmem_alloc<64>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();                      // upstream grids complete; our dependents may start launching
    if (threadIdx.x == 0) C_MARK(1);
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == DG_EPI_WARPS + 1) {
        if (c_elect_one()) {
            c_mbar_expect_tx(wfull, W_BYTES);
            c_tma_3d(&map_w, wfull, w_tile, 0, 0, 0);
            CPatchIter it(tiles_x, tiles_y);
            for (int i = 0; !UNPOOL && i < my_patches; ++i, it.next()) {
                const int s = i % DG_STAGES;
                if (i >= DG_STAGES) c_mbar_wait(&empty[s], ((i / DG_STAGES) - 1) & 1);
                c_mbar_expect_tx(&full[s], Z_HALO_B);
                // halo origin (y0-2, x0-2) of the dz image: pixels outside it are zero-filled by TMA
                c_tma_4d(&map_dz, &full[s], stages + (size_t)s * Z_HALO_B, 0, it.tx * 8 - 2, it.bz * 2, it.ty * 8 - 2);
                if (i == 0) C_MARK(2);
            }
        }
    } else if (warp == DG_EPI_WARPS) {
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(128, CIN, 0, 1);                    // B is MN-major: N = c contiguous
            const uint32_t w0 = c_smem_u32(w_tile);
            c_mbar_wait(wfull, 0);
            for (int i = 0; i < my_patches; ++i) {
                const int s = i % DG_STAGES, b = i & 1;
                c_mbar_wait(&full[s], (i / DG_STAGES) & 1);
                if (UNPOOL) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (i == 0) C_MARK(4);
                if (i >= 2) c_mbar_wait(&tempty[b], ((i >> 1) - 1) & 1);
                c_fence_after();
                const uint32_t a0 = c_smem_u32(stages + (size_t)s * Z_HALO_B);
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
#pragma unroll
                    for (int ks = 0; ks < COUT / 16; ++ks) {
                        // dz[y-kh, x-kw] sits (2-kh, 2-kw) into the halo patch; rows = pixels of 128 B
                        const uint64_t adesc =
                            c_desc(a0 + (2 - kh) * 2 * Z_ROW_B + (2 - kw) * Z_PIX_B + ks * 32, Z_ROW_B, SW128);
                        // W tap [o][32 c]: K = o (rows of 64 B, 8-row groups 512 B apart), N = c inside the row
                        const uint64_t bdesc = c_desc(w0 + t * W_TAP_B + ks * 16 * A_PIX_B, 8 * A_PIX_B, SW64);
                        c_umma(tmem_base + (uint32_t)(b * CIN), adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                c_commit(&empty[s]);
                c_commit(&tfull[b]);
                if (i == 0) C_MARK(5);
                if (i == my_patches - 1) C_MARK(11);
            }
        }
    } else if (warp > DG_EPI_WARPS + 1) {
        // ---------------- un-pool producers (UNPOOL only): 5 x 5 pooled pixels x 2 images x 8 channel groups
        if (UNPOOL) {
            const int ptid = threadIdx.x - (DG_EPI_WARPS + 2) * 32;           // 0..127
            const int PH = (H - 2) / 2, PWp = (W - 2) / 2;
            CPatchIter it(tiles_x, tiles_y);
            for (int i = 0; i < my_patches; ++i, it.next()) {
                const int s = i % DG_STAGES;
                if (i >= DG_STAGES) c_mbar_wait(&empty[s], ((i / DG_STAGES) - 1) & 1);
                uint8_t* tile = stages + (size_t)s * Z_HALO_B;
                // 400 tasks per patch, 4 slots per thread: all loads are issued before the first is consumed
                // (one L2 round trip per patch instead of four)
                CUnpoolIn in[4];
                int row0[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = ptid + u * 128;
                    const int g = t & 7, n = (t >> 3) & 1, pp = min(t >> 4, 24);
                    const int ppy = pp / 5, ppx = pp - ppy * 5;
                    const int py = it.ty * 4 - 1 + ppy, px = it.tx * 4 - 1 + ppx;
                    const bool valid = t < 400 && (unsigned)py < (unsigned)PH && (unsigned)px < (unsigned)PWp;
                    const size_t pidx = ((size_t)(it.bz * 2 + n) * PH + (valid ? py : 0)) * PWp + (valid ? px : 0);
                    in[u] = c_unpool_load(dp, code, valid, pidx, g);
                    row0[u] = (ppy * 2 * 2 + n) * HALO + ppx * 2;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = ptid + u * 128;
                    float gated[8];
                    if (t < 400) c_unpool_emit(tile, in[u], scale, t & 7, row0[u], 2 * HALO, gated);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                c_mbar_arrive(&full[s]);
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const int h = r >> 4, n = (r >> 3) & 1, w = r & 7;
        CPatchIter it(tiles_x, tiles_y);
        for (int i = 0; i < my_patches; ++i, it.next()) {
            const int b = i & 1;
            const int y = it.ty * 8 + h, x = it.tx * 8 + w;
            const bool valid = y < H && x < W;
            const size_t off = (((size_t)(it.bz * 2 + n) * H + y) * W + x) * CIN;
            uint4 g0, g1, g2, g3;
            if (gate && valid) {                                   // issued before the accumulator wait
                g0 = tfy_ld16(gate + off); g1 = tfy_ld16(gate + off + 8);
                g2 = tfy_ld16(gate + off + 16); g3 = tfy_ld16(gate + off + 24);
            }
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
            if (valid) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j >> 4][j & 15]);
                if (gate) {
                    float g[32];
                    TfyPack<__nv_bfloat16>::unpack(g0, g);
                    TfyPack<__nv_bfloat16>::unpack(g1, g + 8);
                    TfyPack<__nv_bfloat16>::unpack(g2, g + 16);
                    TfyPack<__nv_bfloat16>::unpack(g3, g + 24);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = g[j] > 0.f ? v[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) tfy_st16(dx + off + j, TfyPack<__nv_bfloat16>::pack(v + j));
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
// over its share of the patches, stages the fp32 partial tile in shared memory and bulk-copies it to
// `partials[blockIdx.x]` (L2 resident); after a grid-wide arrive/release it reduces a slice of the 18432
// outputs over all partials in a fixed order (deterministic, no atomics) and writes bf16 dW.
// warps 0..15 epilogue + final reduction, warp 16 MMA, warp 17 TMA.
// sync: 1024 uint32, zero-initialised once (generation + two levels of arrival counters, see below).
// ------------------------------------------------------------------------------------------------
// UNPOOL: dz tiles are rebuilt by warps 0..3 from the pooled gradient dp and the forward kernel's code bytes
// (c_unpool_task) and the bias gradient db = sum of the gated dp rides along as 64 more outputs.
template <bool UNPOOL>
__global__ void __launch_bounds__(WG_THREADS, 1)
tfy_conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_a,
                         float* __restrict__ partials, __nv_bfloat16* __restrict__ dw, uint32_t* __restrict__ sync,
                         int tiles_x, int tiles_y, int n_patches, const __nv_bfloat16* __restrict__ dp,
                         const uint8_t* __restrict__ code, float scale, __nv_bfloat16* __restrict__ db, int PH,
                         int PWp, const __grid_constant__ TfyOverlapStep ov) {
    // The last ov.n_cta CTAs of the grid are COMMUNICATION CTAs: they run the fused reduce-scatter -> optimizer ->
    // all-gather step of the parameters whose gradients were final before this kernel started (tfy_fused_step.cuh)
    // over NVLink while the compute CTAs below run the weight gradient.  One CTA per SM and grid <= #SMs, so both
    // roles are co-resident by construction.
    const int n_cmp = (int)gridDim.x - ov.n_cta;                     // compute CTAs (share patches, grid barrier)
    if ((int)blockIdx.x >= n_cmp) {
        extern __shared__ uint8_t smem_role[];
        tfy_pdl_sync();
        tfy_overlap_role<WG_THREADS>(ov, (int)blockIdx.x - n_cmp, c_align1024(smem_role));
        return;
    }
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align1024(smem_raw);        // [4] { dz [8 h][2 n][8 w][128 B] sw128 ; a [10 h][2 n][10 w][64 B] sw64 }
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)WG_STAGES * WG_STAGE_B);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tap_full = empty + WG_STAGES;                         // [9]: accumulator of tap t is final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tap_full + TAPS);
    float* s_db = reinterpret_cast<float*>(tmem_slot + 4);          // [16 warps][64] bias-gradient partials (UNPOOL)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        if (!UNPOOL) c_prefetch_map(&map_dz);
        c_prefetch_map(&map_a);
        for (int s = 0; s < WG_STAGES; ++s) { c_mbar_init(&full[s], UNPOOL ? 257 : 1); c_mbar_init(&empty[s], 1); }
        for (int t = 0; t < TAPS; ++t) c_mbar_init(&tap_full[t], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WG_WORK_WARPS) c_tmem_alloc<512>(tmem_slot);
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();                      // upstream grids complete; our dependents may start launching
    if (threadIdx.x == 0) C_MARK(1);
    const uint32_t gen0 = threadIdx.x == 0 ? *reinterpret_cast<volatile uint32_t*>(sync) : 0u;
    const int my_patches = ((int)blockIdx.x < n_patches) ? (n_patches - 1 - (int)blockIdx.x) / n_cmp + 1 : 0;

    if (warp == WG_WORK_WARPS + 1) {
        if (c_elect_one()) {
            CPatchIter it(tiles_x, tiles_y, n_cmp);
            for (int i = 0; i < my_patches; ++i, it.next()) {
                const int s = i % WG_STAGES;
                if (i >= WG_STAGES) c_mbar_wait(&empty[s], ((i / WG_STAGES) - 1) & 1);
                uint8_t* st = stages + (size_t)s * WG_STAGE_B;
                c_mbar_expect_tx(&full[s], (UNPOOL ? 0 : Z_PATCH_B) + A_HALO_B);
                if (!UNPOOL) c_tma_4d(&map_dz, &full[s], st, 0, it.tx * 8, it.bz * 2, it.ty * 8);
                c_tma_4d(&map_a, &full[s], st + Z_PATCH_B, 0, it.tx * 8, it.bz * 2, it.ty * 8);
                if (i == 0) C_MARK(2);
            }
        }
    } else if (warp == WG_WORK_WARPS) {
        if (c_elect_one()) {
            // Patches are taken WG_STAGES at a time and the taps form the OUTER loop: the accumulator of tap t
            // is final after the last group's pass over t, so its store overlaps the MMAs of taps t+1..8.
            const uint32_t idesc = c_idesc(64, CIN, 1, 1);
            const int n_groups = (my_patches + WG_STAGES - 1) / WG_STAGES;
            for (int g = 0; g < n_groups; ++g) {
                const int cnt = min(WG_STAGES, my_patches - g * WG_STAGES);
                for (int q = 0; q < cnt; ++q) c_mbar_wait(&full[q], g & 1);
                if (UNPOOL) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (g == 0) C_MARK(4);
                c_fence_after();
                const uint32_t st0 = c_smem_u32(stages);
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
                    const int kh = t / 3, kw = t - kh * 3;
#pragma unroll 1
                    for (int q = 0; q < cnt; ++q) {
                        const uint32_t dz0 = st0 + q * WG_STAGE_B, a0 = dz0 + Z_PATCH_B;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {      // 16 pixels (one h row of both images) per MMA
                            // A = dz^T: M = 64 o inside the 128 B pixel row; K = pixels, 8-pixel groups 1024 B apart
                            const uint64_t adesc = c_desc(dz0 + j * 2 * 8 * Z_PIX_B, 8 * Z_PIX_B, SW128);
                            // B = a (tap shifted): N = 32 c inside the 64 B pixel row; K = pixels, groups = (h, n) rows
                            const uint64_t bdesc = c_desc(a0 + (j + kh) * 2 * A_ROW_B + kw * A_PIX_B, A_ROW_B, SW64);
                            c_umma(tmem_base + (uint32_t)(t * CIN), adesc, bdesc, idesc, (g > 0 || q > 0 || j > 0) ? 1u : 0u);
                        }
                    }
                    if (g == n_groups - 1) c_commit(&tap_full[t]);
                }
                for (int q = 0; q < cnt; ++q) c_commit(&empty[q]);     // stages reusable once these MMAs retire
            }
            C_MARK(5);
        }
    } else if (my_patches > 0) {
        if (UNPOOL) {
            // ---------------- un-pool producers (all 16 worker warps): a patch needs 4 x 4 pooled pixels x 2 images
            // x 8 channel groups = 256 tasks; thread t serves task t & 255 of patches (t >> 8), (t >> 8) + 2, ...
            // and always the same channel group, so db accumulates in registers.  Loads of two patches are in
            // flight together.
            const int t = threadIdx.x, task = t & 255, g = task & 7, n = (task >> 3) & 1, pp = task >> 4;
            const int ppy = pp >> 2, ppx = pp & 3;
            const int row0 = (ppy * 2 * 2 + n) * 8 + ppx * 2;
            float colacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            CPatchIter it(tiles_x, tiles_y, n_cmp);
            if (t >> 8) it.next();
            for (int i = (t >> 8); i < my_patches; i += 4) {
                // patches i and i + 2 (if any)
                CPatchIter it2 = it;
                it2.next(); it2.next();
                const bool two = i + 2 < my_patches;
                const size_t p0 = ((size_t)(it.bz * 2 + n) * PH + it.ty * 4 + ppy) * PWp + it.tx * 4 + ppx;
                const size_t p1 = ((size_t)(it2.bz * 2 + n) * PH + it2.ty * 4 + ppy) * PWp + it2.tx * 4 + ppx;
                const CUnpoolIn in0 = c_unpool_load(dp, code, true, p0, g);
                const CUnpoolIn in1 = c_unpool_load(dp, code, two, two ? p1 : p0, g);
                float gated[8];
                {
                    const int s = i % WG_STAGES;
                    if (i >= WG_STAGES) c_mbar_wait(&empty[s], ((i / WG_STAGES) - 1) & 1);
                    c_unpool_emit(stages + (size_t)s * WG_STAGE_B, in0, scale, g, row0, 16, gated);
#pragma unroll
                    for (int k = 0; k < 8; ++k) colacc[k] += gated[k];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    c_mbar_arrive(&full[s]);
                }
                if (two) {
                    const int i2 = i + 2, s = i2 % WG_STAGES;
                    if (i2 >= WG_STAGES) c_mbar_wait(&empty[s], ((i2 / WG_STAGES) - 1) & 1);
                    c_unpool_emit(stages + (size_t)s * WG_STAGE_B, in1, scale, g, row0, 16, gated);
#pragma unroll
                    for (int k = 0; k < 8; ++k) colacc[k] += gated[k];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    c_mbar_arrive(&full[s]);
                }
                it.next(); it.next(); it.next(); it.next();
            }
            // lanes l, l+8, l+16, l+24 serve the same channel group: reduce them, one plain store per warp
            // (shared-memory float atomics are CAS loops: 512 threads on 64 addresses cost microseconds)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                colacc[k] += __shfl_xor_sync(0xffffffffu, colacc[k], 8);
                colacc[k] += __shfl_xor_sync(0xffffffffu, colacc[k], 16);
            }
            if (lane < 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) s_db[warp * COUT + g * 8 + k] = colacc[k];
            }
        }
        // M = 64 accumulator: row o lives in TMEM lane (o % 16) + 32 * (o / 16): lanes 0..15 of each quadrant.
        // warp = (quadrant, tap group): 16 lanes write the 128-byte (o, tap) rows of the fp32 partial tile.
        const int quad = warp & 3, tg = warp >> 2;                       // taps tg, tg+4, tg+8
        float* mine = partials + (size_t)blockIdx.x * WG_PART + (size_t)(quad * 16 + (lane & 15)) * (TAPS * CIN);
#pragma unroll 1
        for (int t = tg; t < TAPS; t += 4) {
            c_mbar_wait(&tap_full[t], 0);
            if (threadIdx.x == 0 && t == 0) C_MARK(6);
            c_fence_after();
            uint32_t acc[2][16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * CIN);
            c_tmem_ld16(taddr, acc[0]);
            c_tmem_ld16(taddr + 16, acc[1]);
            c_tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    __stcg(reinterpret_cast<float4*>(mine + t * CIN + j),
                           make_float4(__uint_as_float(acc[j >> 4][j & 15]), __uint_as_float(acc[j >> 4][(j & 15) + 1]),
                                       __uint_as_float(acc[j >> 4][(j & 15) + 2]),
                                       __uint_as_float(acc[j >> 4][(j & 15) + 3])));
            }
        }
        if (UNPOOL) {
            asm volatile("bar.sync 1, %0;" ::"r"(WG_WORK_WARPS * 32) : "memory");     // all producers added their db
            if (threadIdx.x < COUT) {
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < WG_WORK_WARPS; ++q) a += s_db[q * COUT + threadIdx.x];
                __stcg(partials + (size_t)blockIdx.x * WG_PART + WG_OUT + threadIdx.x, a);
            }
        }
        __threadfence();
    }
    c_fence_before();
    __syncthreads();
    if (warp == WG_WORK_WARPS) c_tmem_free<512>(tmem_base);
    if (threadIdx.x == 0) C_MARK(7);

    // Grid-wide arrive / release (all CTAs are co-resident: grid <= #SMs, one CTA per SM).  Same-sector atomics
    // serialise at ~20-30 cycles each and every returning atomic in a chain costs an L2 round trip, so: arrivals
    // are fire-and-forget `red`s spread over 16 counters (one per 128-byte line, sync[32 + 32 k]); CTA 0 polls
    // them, clears them and bumps the generation word sync[0] that everybody else polls.
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync + 32 + 32 * (blockIdx.x & 15u)) : "memory");
    }
    if (blockIdx.x == 0 && warp == 0) {
        if (lane < 16 && lane < n_cmp) {
            const uint32_t expect = ((uint32_t)n_cmp - (uint32_t)lane + 15u) / 16u;
            volatile uint32_t* cnt = sync + 32 + 32 * lane;
            while (*cnt != expect) __nanosleep(20);
            *cnt = 0u;
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(sync) = gen0 + 1;
        }
    }
    if (threadIdx.x == 0) {
        while (*reinterpret_cast<volatile uint32_t*>(sync) == gen0) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) C_MARK(8);
    // slice of the outputs owned by this CTA (float4 units), summed over the partials in a fixed order
    constexpr int TOTAL4 = (UNPOOL ? WG_PART : WG_OUT) / 4, STRIDE4 = WG_PART / 4;
    const int per = (TOTAL4 + n_cmp - 1) / n_cmp;
    const int lo = (int)blockIdx.x * per, hi = min(lo + per, TOTAL4);
    const int len = max(hi - lo, 0);
    const int n_part = n_cmp;                                       // every CTA wrote a partial
    const float4* src = reinterpret_cast<const float4*>(partials) + lo;
    auto store = [&](int f, const float4& s4) {
        __nv_bfloat162 l2 = __floats2bfloat162_rn(s4.x, s4.y), h2 = __floats2bfloat162_rn(s4.z, s4.w);
        uint2 packed;
        packed.x = *reinterpret_cast<uint32_t*>(&l2);
        packed.y = *reinterpret_cast<uint32_t*>(&h2);
        const int e = (lo + f) * 4;                                  // outputs [0, WG_OUT) = dW, then db
        if (e < WG_OUT) *reinterpret_cast<uint2*>(dw + e) = packed;
        else *reinterpret_cast<uint2*>(db + (e - WG_OUT)) = packed;
    };
    if (len > 0 && 2 * len <= WG_THREADS) {
        // few outputs per CTA (the usual case: 32): `parts` thread groups split the list of partials
        float4* red = reinterpret_cast<float4*>(stages);             // the ring is idle now
        const int parts = WG_THREADS / len;
        const int f = (int)threadIdx.x % len, part = (int)threadIdx.x / len;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (part < parts) {
#pragma unroll 8
            for (int q = part; q < n_part; q += parts) {
                const float4 v = __ldcg(src + (size_t)q * STRIDE4 + f);
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
                const float4 v = __ldcg(src + (size_t)q * STRIDE4 + f);
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
            store(f, s4);
        }
    }
    if (threadIdx.x == 0) C_MARK(10);
}

// ------------------------------------------------------------------------------------------------
// First-layer weight + bias gradient (C_in = 1, 32 filters) on the tensor core:
//   D[n, o] = sum_pix X~[pix, n] dz[pix, o],   X~[pix, 0..8] = x[pix + tap], X~[pix, 9] = 1 (bias), rest 0
// M = 64 (n, padded), N = 32 (o), K = pixels.  dz rows (64 B) arrive by TMA exactly as they sit in memory
// (MN-major B, 64B swizzle); the im2col operand X~ (128-byte rows, 128B swizzle, MN-major A) is built in
// shared memory by 128 threads from the tiny single-channel input.  Replaces the CUDA-core wgrad kernel
// (28.8 us) and the separate bias-gradient column sum (8.2 us) of the first layer.
// warps 0..3 build X~ (warp 0 also runs the epilogue), warp 4 MMA, warp 5 TMA.
// acc: 16 replicas of 320 floats [n][o] (same-sector atomics serialise, so CTAs spread over replicas), zero on
// entry, left zero; counter: one uint32, zero on entry, left zero.
// ------------------------------------------------------------------------------------------------
constexpr int C1_O = 32, C1_STAGES = 4, C1_PIX = 128, C1_REPLICAS = 16;
constexpr int C1_A_B = C1_PIX * 128, C1_B_B = C1_PIX * 64, C1_STAGE_B = C1_A_B + C1_B_B;   // 16 KB + 8 KB
constexpr int C1_THREADS = 192;
constexpr size_t C1_SMEM = (size_t)C1_STAGES * C1_STAGE_B + 1024 + 256;

template <typename XT>
__global__ void __launch_bounds__(C1_THREADS, 1)
tfy_conv3x3_c1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz, const XT* __restrict__ x,
                               float* __restrict__ acc, uint32_t* __restrict__ counter,
                               __nv_bfloat16* __restrict__ dw, __nv_bfloat16* __restrict__ dbias, int H, int W,
                               int n_pix, int n_chunks) {
    extern __shared__ uint8_t smem_raw[];
    C_TIMELINE_BEGIN();
    uint8_t* stages = c_align1024(smem_raw);        // [4] { X~ [128 pix][128 B] sw128 ; dz [128 pix][64 B] sw64 }
    uint64_t* full_a = reinterpret_cast<uint64_t*>(stages + (size_t)C1_STAGES * C1_STAGE_B);
    uint64_t* full_b = full_a + C1_STAGES;
    uint64_t* empty = full_b + C1_STAGES;
    uint64_t* tmem_full = empty + C1_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    __shared__ uint32_t s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        c_prefetch_map(&map_dz);
        for (int s = 0; s < C1_STAGES; ++s) {
            c_mbar_init(&full_a[s], 128);
            c_mbar_init(&full_b[s], 1);
            c_mbar_init(&empty[s], 1);
        }
        c_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) c_tmem_alloc<32>(tmem_slot);
    // the padding columns of X~ (n >= 16) stay zero for the whole kernel
    for (int i = threadIdx.x; i < C1_STAGES * C1_A_B / 16; i += C1_THREADS) {
        const int s = i / (C1_A_B / 16), r = i - s * (C1_A_B / 16);
        *reinterpret_cast<uint4*>(stages + (size_t)s * C1_STAGE_B + (size_t)r * 16) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    tfy_pdl_sync();                      // upstream grids complete; our dependents may start launching
    if (threadIdx.x == 0) C_MARK(1);
    const int my_chunks = ((int)blockIdx.x < n_chunks) ? (n_chunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int OH = H - 2, OW = W - 2;

    if (warp == 5) {
        if (c_elect_one()) {
            for (int i = 0; i < my_chunks; ++i) {
                const int s = i % C1_STAGES;
                if (i >= C1_STAGES) c_mbar_wait(&empty[s], ((i / C1_STAGES) - 1) & 1);
                const int chunk = (int)blockIdx.x + i * (int)gridDim.x;
                c_mbar_expect_tx(&full_b[s], C1_B_B);
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                    ::"r"(c_smem_u32(stages + (size_t)s * C1_STAGE_B + C1_A_B)),
                    "l"(reinterpret_cast<uint64_t>(&map_dz)), "r"(c_smem_u32(&full_b[s])), "r"(0), "r"(chunk * C1_PIX)
                    : "memory");
            }
        }
    } else if (warp == 4) {
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(64, C1_O, 1, 1);
            for (int i = 0; i < my_chunks; ++i) {
                const int s = i % C1_STAGES;
                c_mbar_wait(&full_a[s], (i / C1_STAGES) & 1);
                c_mbar_wait(&full_b[s], (i / C1_STAGES) & 1);
                if (i == 0) C_MARK(4);
                c_fence_after();
                const uint32_t a0 = c_smem_u32(stages + (size_t)s * C1_STAGE_B), b0 = a0 + C1_A_B;
#pragma unroll
                for (int j = 0; j < C1_PIX / 16; ++j) {
                    // A = X~^T: M = 64 n inside the 128 B pixel row; B = dz: N = 32 o inside the 64 B pixel row
                    const uint64_t adesc = c_desc(a0 + j * 16 * 128, 8 * 128, SW128);
                    const uint64_t bdesc = c_desc(b0 + j * 16 * 64, 8 * 64, SW64);
                    c_umma(tmem_base, adesc, bdesc, idesc, (i > 0 || j > 0) ? 1u : 0u);
                }
                c_commit(&empty[s]);
            }
            c_commit(tmem_full);
            C_MARK(5);
        }
    } else {
        // ---------------- im2col builders: one pixel per thread per chunk; the next chunk's inputs are gathered
        // before this chunk is packed and stored (one exposed L2 round trip instead of one per chunk)
        const int r = threadIdx.x;                                  // row of the stage, 0..127
        const uint32_t sw = (uint32_t)(r & 7);
        auto gather = [&](int chunk, float* v) -> bool {
            const int p = chunk * C1_PIX + r;
            const bool ok = p < n_pix;
            if (ok) {
                const int b = p / (OH * OW), rem = p - b * (OH * OW), y = rem / OW, xx = rem - y * OW;
                const XT* src = x + ((size_t)b * H + y) * W + xx;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) v[kh * 3 + kw] = (float)src[kh * W + kw];
            } else {
#pragma unroll
                for (int t = 0; t < 9; ++t) v[t] = 0.f;
            }
            return ok;
        };
        float v[9];
        bool ok = my_chunks > 0 ? gather((int)blockIdx.x, v) : false;
        for (int i = 0; i < my_chunks; ++i) {
            const int s = i % C1_STAGES;
            float c1[8] = {v[8], ok ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const uint4 q0 = TfyPack<__nv_bfloat16>::pack(v), q1 = TfyPack<__nv_bfloat16>::pack(c1);
            if (i + 1 < my_chunks) ok = gather((int)blockIdx.x + (i + 1) * (int)gridDim.x, v);
            if (i >= C1_STAGES) c_mbar_wait(&empty[s], ((i / C1_STAGES) - 1) & 1);
            uint8_t* rowp = stages + (size_t)s * C1_STAGE_B + (size_t)r * 128;
            *reinterpret_cast<uint4*>(rowp + ((0u ^ sw) << 4)) = q0;    // 128B swizzle: chunk ^= row & 7
            *reinterpret_cast<uint4*>(rowp + ((1u ^ sw) << 4)) = q1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            c_mbar_arrive(&full_a[s]);
            if (threadIdx.x == 0) { if (i == 0) C_MARK(2); if (i == my_chunks - 1) C_MARK(3); }
        }
        if (warp == 0 && my_chunks > 0) {
            // D row n lives in TMEM lane n (n < 16 -> quadrant 0); lanes 0..8 = taps, lane 9 = bias
            c_mbar_wait(tmem_full, 0);
            if (threadIdx.x == 0) C_MARK(6);
            c_fence_after();
            uint32_t d[2][16];
            c_tmem_ld16(tmem_base, d[0]);
            c_tmem_ld16(tmem_base + 16, d[1]);
            c_tmem_ld_wait();
            if (lane < 10) {
#pragma unroll
                for (int o = 0; o < C1_O; ++o)
                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(acc + (blockIdx.x % C1_REPLICAS) * (10 * C1_O) +
                                                                        lane * C1_O + o),
                                 "f"(__uint_as_float(d[o >> 4][o & 15]))
                                 : "memory");
            }
            __threadfence();
            if (threadIdx.x == 0) C_MARK(7);
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == 4) c_tmem_free<32>(tmem_base);
    // the last CTA to finish converts the sums and clears the accumulator for the next step
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) C_MARK(8);
    if (s_last) {
        __threadfence();
        for (int i = threadIdx.x; i < 10 * C1_O; i += C1_THREADS) {
            const int n = i / C1_O, o = i - n * C1_O;
            float val = 0.f;
#pragma unroll
            for (int rep = 0; rep < C1_REPLICAS; ++rep) {
                val += __ldcg(acc + rep * (10 * C1_O) + i);
                __stcg(acc + rep * (10 * C1_O) + i, 0.f);
            }
            if (n < 9) dw[o * 9 + n] = __float2bfloat16(val);
            else dbias[o] = __float2bfloat16(val);
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
    if (threadIdx.x == 0) C_MARK(10);
}

// ------------------------------------------------------------------------------------------------
// First-layer forward (C_in = 1, 32 filters) on the tensor core:
//   y[pix, o] = relu( sum_k X~[pix, k] W~[o, k] ),  X~[pix, 0..8] = x[pix + tap], X~[pix, 9] = 1,  W~[o, 9] = bias[o]
// One M128 N32 K16 tcgen05.mma per 128 pixels; both operands are built in shared memory in the un-swizzled
// K-major core-matrix layout ([k half][row][16 B]: lbo = distance between the halves, sbo = 128).  The CUDA-core
// version spent its time on 288 shared-memory weight loads + FMAs per pixel; here a thread gathers 9 inputs,
// stores 2 x 16 B, and later packs 32 accumulator columns.
// warps 0..3 build X~, warps 4..7 epilogue (TMEM quadrant = warp & 3), warp 8 MMA.
// ------------------------------------------------------------------------------------------------
constexpr int F1_STAGES = 4, F1_PIX = 128, F1_A_B = 2 * F1_PIX * 16;      // 4 KB per stage
constexpr int F1_W_B = 2 * C1_O * 16;                                     // 1 KB
constexpr int F1_THREADS = 9 * 32;
constexpr size_t F1_SMEM = (size_t)F1_STAGES * F1_A_B + F1_W_B + 128 + 256;

__device__ __forceinline__ uint64_t c_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <typename XT>
__global__ void __launch_bounds__(F1_THREADS, 2)
tfy_conv3x3_c1_fwd_tc_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                             const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y, int H, int W,
                             int n_pix, int n_chunks) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* w_tile = stages + (size_t)F1_STAGES * F1_A_B;          // [2 k halves][32 o][16 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(w_tile + F1_W_B);
    uint64_t* empty = full + F1_STAGES;
    uint64_t* tfull = empty + F1_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < F1_STAGES; ++s) { c_mbar_init(&full[s], 128); c_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { c_mbar_init(&tfull[b], 1); c_mbar_init(&tempty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) c_tmem_alloc<64>(tmem_slot);
    tfy_pdl_sync();                      // (the weights below are produced by the previous step's optimizer kernel)
    if (threadIdx.x < C1_O) {
        // W~ row o: taps 0..7 | tap 8, bias, 0 x 6
        const int o = threadIdx.x;
        float k0[8], k1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 8; ++t) k0[t] = __bfloat162float(w[o * 9 + t]);
        k1[0] = __bfloat162float(w[o * 9 + 8]);
        k1[1] = __bfloat162float(bias[o]);
        *reinterpret_cast<uint4*>(w_tile + o * 16) = TfyPack<__nv_bfloat16>::pack(k0);
        *reinterpret_cast<uint4*>(w_tile + C1_O * 16 + o * 16) = TfyPack<__nv_bfloat16>::pack(k1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    c_fence_before();
    __syncthreads();
    c_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_chunks = ((int)blockIdx.x < n_chunks) ? (n_chunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int OH = H - 2, OW = W - 2;

    if (warp == 8) {
        if (c_elect_one()) {
            const uint32_t idesc = c_idesc(128, C1_O, 0, 0);
            const uint64_t bdesc = c_desc_nosw(c_smem_u32(w_tile), C1_O * 16, 128);
            for (int i = 0; i < my_chunks; ++i) {
                const int s = i % F1_STAGES, b = i & 1;
                c_mbar_wait(&full[s], (i / F1_STAGES) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (i >= 2) c_mbar_wait(&tempty[b], ((i >> 1) - 1) & 1);
                c_fence_after();
                const uint64_t adesc = c_desc_nosw(c_smem_u32(stages + (size_t)s * F1_A_B), F1_PIX * 16, 128);
                c_umma(tmem_base + (uint32_t)(b * C1_O), adesc, bdesc, idesc, 0u);
                c_commit(&empty[s]);
                c_commit(&tfull[b]);
            }
        }
    } else if (warp < 4) {
        // ---------------- im2col builders: one pixel per thread per chunk; the next chunk's inputs are loaded
        // before this chunk's stage is waited for
        const int r = threadIdx.x;
        auto gather = [&](int chunk, float* v) -> bool {
            const int p = chunk * F1_PIX + r;
            const bool ok = p < n_pix;
            if (ok) {
                const int b = p / (OH * OW), rem = p - b * (OH * OW), yy = rem / OW, xx = rem - yy * OW;
                const XT* src = x + ((size_t)b * H + yy) * W + xx;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) v[kh * 3 + kw] = (float)src[kh * W + kw];
            } else {
#pragma unroll
                for (int t = 0; t < 9; ++t) v[t] = 0.f;
            }
            return ok;
        };
        float v[9];
        bool ok = my_chunks > 0 ? gather((int)blockIdx.x, v) : false;
        for (int i = 0; i < my_chunks; ++i) {
            const int s = i % F1_STAGES;
            float c1[8] = {v[8], ok ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const uint4 q0 = TfyPack<__nv_bfloat16>::pack(v), q1 = TfyPack<__nv_bfloat16>::pack(c1);
            if (i + 1 < my_chunks) ok = gather((int)blockIdx.x + (i + 1) * (int)gridDim.x, v);
            if (i >= F1_STAGES) c_mbar_wait(&empty[s], ((i / F1_STAGES) - 1) & 1);
            uint8_t* st = stages + (size_t)s * F1_A_B;
            *reinterpret_cast<uint4*>(st + r * 16) = q0;
            *reinterpret_cast<uint4*>(st + F1_PIX * 16 + r * 16) = q1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            c_mbar_arrive(&full[s]);
        }
    } else {
        // ---------------- epilogue: accumulator row = pixel of the chunk; ReLU, bf16, 64 contiguous bytes
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        for (int i = 0; i < my_chunks; ++i) {
            const int b = i & 1;
            const int chunk = (int)blockIdx.x + i * (int)gridDim.x;
            c_mbar_wait(&tfull[b], (i >> 1) & 1);
            c_fence_after();
            uint32_t acc[2][16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * C1_O);
            c_tmem_ld16(taddr, acc[0]);
            c_tmem_ld16(taddr + 16, acc[1]);
            c_tmem_ld_wait();
            c_fence_before();
            c_mbar_arrive(&tempty[b]);
            const int p = chunk * F1_PIX + r;
            if (p < n_pix) {
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaxf(__uint_as_float(acc[j >> 4][j & 15]), 0.f);
                __nv_bfloat16* yp = y + (size_t)p * C1_O;
#pragma unroll
                for (int j = 0; j < 32; j += 8) tfy_st16(yp + j, TfyPack<__nv_bfloat16>::pack(o + j));
            }
        }
    }
    c_fence_before();
    __syncthreads();
    if (warp == 8) c_tmem_free<64>(tmem_base);
}

namespace {
using CEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
CEncodeFn c_encode = nullptr;
bool c_attr_set = false;
int c_sms = 0;
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
                                 (int)FP_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)DG_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)DG_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)WG_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
// NHWC bf16 tensor [B, H, W, C] seen as {C, W, B, H}; box = {C, bw, 2, bh}: one pixel per shared-memory row
bool c_map_nhwc(CUtensorMap* m, const void* p, int B, int H, int W, int C, int bw, int bh) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, 2, (cuuint32_t)bh};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return c_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weights [O][9][C] bf16 seen as {C, O, 9}; one box = everything, stored [tap][o][32 c] with 64-byte rows
bool c_map_w(CUtensorMap* m, const void* p) {
    cuuint64_t dims[3] = {(cuuint64_t)CIN, (cuuint64_t)COUT, (cuuint64_t)TAPS};
    cuuint64_t strides[2] = {(cuuint64_t)TAPS * CIN * 2, (cuuint64_t)CIN * 2};
    cuuint32_t box[3] = {(cuuint32_t)CIN, (cuuint32_t)COUT, (cuuint32_t)TAPS};
    cuuint32_t estr[3] = {1, 1, 1};
    return c_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" {

// debug: device buffer of 16 x int64 per CTA receiving the event timeline of the next launches (nullptr = off)
int tfy_conv_set_timeline(long long* buf) {
    return (int)cudaMemcpyToSymbol(c_timeline, &buf, sizeof(buf));
}

// number of fp32 elements of the `partials` scratch buffer tfy_conv3x3_c32_wgrad needs
size_t tfy_conv3x3_c32_wgrad_scratch_elems() { return (size_t)160 * WG_PART; }

// a: [B, H, W, 32] bf16 (NHWC), w: [64, 3, 3, 32] bf16, bias: [64] bf16
// pooled / code: [B, (H-2)/2, (W-2)/2, 64].  Requires (H-2) % 8 == 0, (W-2) % 8 == 0, B % 2 == 0.
int tfy_conv3x3_c32_pool_fwd(const void* a, const void* w, const void* bias, void* pooled, void* code, int B, int H,
                             int W, float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2) || drop_rate < 0.f || drop_rate >= 1.f) return -2;
    if ((size_t)B * (OH / 2) * (OW / 2) * COUT >= (1ull << 32)) return -3;
    if (!c_init()) return -4;
    CUtensorMap ma, mw;
    if (!c_map_nhwc(&ma, a, B, H, W, CIN, HALO, HALO) || !c_map_w(&mw, w)) return -6;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_launch_pdl((tfy_conv3x3_fprop_pool_kernel), dim3(c_grid(n_patches)), dim3(FP_THREADS), FP_SMEM, s, 
        ma, mw, (const __nv_bfloat16*)bias, (__nv_bfloat16*)pooled, (uint8_t*)code, H, W, tiles_x, tiles_y, n_patches,
        drop_rate, seed, hp);
    return (int)cudaGetLastError();
}

// dz: [B, H-2, W-2, 64] bf16, w: [64, 3, 3, 32], gate (optional): [B, H, W, 32], dx: [B, H, W, 32].  B % 2 == 0.
int tfy_conv3x3_c32_dgrad(const void* dz, const void* w, const void* gate, void* dx, int B, int H, int W,
                          cudaStream_t s) {
    if (B % 2) return -2;
    if (!c_init()) return -4;
    CUtensorMap mz, mw;
    if (!c_map_nhwc(&mz, dz, B, H - 2, W - 2, COUT, HALO, HALO) || !c_map_w(&mw, w)) return -6;
    const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_launch_pdl((tfy_conv3x3_dgrad_kernel<false>), dim3(c_grid(n_patches)), dim3(DG_THREADS), DG_SMEM, s, 
        mz, mw, (const __nv_bfloat16*)gate, (__nv_bfloat16*)dx, H, W, tiles_x, tiles_y, n_patches, (const __nv_bfloat16*)nullptr, (const uint8_t*)nullptr, 1.0f);
    return (int)cudaGetLastError();
}

// a: [B, H, W, 32], dz: [B, H-2, W-2, 64], dw: [64, 3, 3, 32] bf16 (overwritten).
// partials: tfy_conv3x3_c32_wgrad_scratch_elems() floats of scratch; sync: 1024 x uint32, zeroed once at allocation.
int tfy_conv3x3_c32_wgrad(const void* a, const void* dz, float* partials, void* dw, uint32_t* sync, int B, int H, int W,
                          cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2)) return -2;
    if (!c_init() || c_sms > 160) return -4;
    CUtensorMap mz, ma;
    if (!c_map_nhwc(&mz, dz, B, OH, OW, COUT, 8, 8) || !c_map_nhwc(&ma, a, B, H, W, CIN, HALO, HALO)) return -6;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_launch_pdl((tfy_conv3x3_wgrad_kernel<false>), dim3(c_grid(n_patches)), dim3(WG_THREADS), WG_SMEM, s, mz, ma,
                   partials, (__nv_bfloat16*)dw, sync, tiles_x, tiles_y, n_patches, (const __nv_bfloat16*)nullptr,
                   (const uint8_t*)nullptr, 1.0f, (__nv_bfloat16*)nullptr, 0, 0, TfyOverlapStep{});
    return (int)cudaGetLastError();
}

// Fused pool/dropout/ReLU backward variants: instead of a materialised dz they take the pooled gradient
// dp [B, (H-2)/2, (W-2)/2, 64] bf16, the forward kernel's code bytes (same shape) and the dropout scale.
int tfy_conv3x3_c32_dgrad_unpool(const void* dp, const void* code, float scale, const void* w, const void* gate,
                                 void* dx, int B, int H, int W, cudaStream_t s) {
    if ((B % 2) || ((H - 2) % 8) || ((W - 2) % 8)) return -2;
    if (!c_init()) return -4;
    CUtensorMap mw;
    if (!c_map_w(&mw, w)) return -6;
    const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8, n_patches = tiles_x * tiles_y * (B / 2);
    tfy_launch_pdl((tfy_conv3x3_dgrad_kernel<true>), dim3(c_grid(n_patches)), dim3(DG_THREADS), DG_SMEM, s, mw, mw,
                   (const __nv_bfloat16*)gate, (__nv_bfloat16*)dx, H, W, tiles_x, tiles_y, n_patches,
                   (const __nv_bfloat16*)dp, (const uint8_t*)code, scale);
    return (int)cudaGetLastError();
}

// also writes db [64] = sum over batch and positions of the gated pooled gradient.
// ov (optional): a fused gradient-exchange/optimizer step to run on SPARE SMs next to the weight gradient
// (tfy_overlap_role).  ov->n_cta is clamped to the SMs the compute grid leaves free (576 patches -> 144 compute
// CTAs on a 148-SM B200 -> 4 communication CTAs); tfy_conv3x3_c32_wgrad_spare_ctas() tells the caller how many
// that is so that it can size the overlapped range.
int tfy_conv3x3_c32_wgrad_unpool_ov(const void* a, const void* dp, const void* code, float scale, float* partials,
                                    void* dw, void* db, uint32_t* sync, int B, int H, int W, const TfyOverlapStep* ov,
                                    cudaStream_t s) {
    const int OH = H - 2, OW = W - 2;
    if ((OH % 8) || (OW % 8) || (B % 2)) return -2;
    if (!c_init() || c_sms > 160) return -4;
    CUtensorMap ma;
    if (!c_map_nhwc(&ma, a, B, H, W, CIN, HALO, HALO)) return -6;
    const int tiles_x = OW / 8, tiles_y = OH / 8, n_patches = tiles_x * tiles_y * (B / 2);
    const int n_cmp = c_grid(n_patches);
    TfyOverlapStep o{};
    if (ov != nullptr && ov->n_cta > 0) {
        o = *ov;
        if (o.n_cta > c_sms - n_cmp) return -7;       // the role must not displace a compute CTA
        if (o.slot0 < 0 || o.slot0 + o.n_cta > TFY_MAX_BLOCKS) return -8;
    }
    tfy_launch_pdl((tfy_conv3x3_wgrad_kernel<true>), dim3(n_cmp + o.n_cta), dim3(WG_THREADS), WG_SMEM, s, ma, ma,
                   partials, (__nv_bfloat16*)dw, sync, tiles_x, tiles_y, n_patches, (const __nv_bfloat16*)dp,
                   (const uint8_t*)code, scale, (__nv_bfloat16*)db, OH / 2, OW / 2, o);
    return (int)cudaGetLastError();
}

int tfy_conv3x3_c32_wgrad_unpool(const void* a, const void* dp, const void* code, float scale, float* partials,
                                 void* dw, void* db, uint32_t* sync, int B, int H, int W, cudaStream_t s) {
    return tfy_conv3x3_c32_wgrad_unpool_ov(a, dp, code, scale, partials, dw, db, sync, B, H, W, nullptr, s);
}

// SMs the weight-gradient grid leaves idle for B images of H x W (0 when the patches fill every SM)
int tfy_conv3x3_c32_wgrad_spare_ctas(int B, int H, int W) {
    if (!c_init()) return 0;
    const int n_patches = ((W - 2) / 8) * ((H - 2) / 8) * (B / 2);
    const int spare = c_sms - c_grid(n_patches);
    return spare > 0 ? spare : 0;
}

// First-layer (C_in = 1, 32 filters) forward with bias + ReLU.  x: [B, H, W, 1] fp32 or bf16, w: [32, 3, 3, 1] bf16,
// bias: [32] bf16, y: [B, H-2, W-2, 32] bf16.
int tfy_conv3x3_c1_fwd_tc(const void* x, int x_is_f32, const void* w, const void* bias, void* y, int B, int H, int W,
                          cudaStream_t s) {
    if (!c_init()) return -4;
    const long long n_pix_ll = (long long)B * (H - 2) * (W - 2);
    if (n_pix_ll >= (1ll << 31) - F1_PIX) return -3;
    const int n_pix = (int)n_pix_ll, n_chunks = (n_pix + F1_PIX - 1) / F1_PIX;
    const int grid = n_chunks < 2 * c_sms ? n_chunks : 2 * c_sms;
    if (x_is_f32)
        tfy_launch_pdl((tfy_conv3x3_c1_fwd_tc_kernel<float>), dim3(grid), dim3(F1_THREADS), F1_SMEM, s, (const float*)x,
                       (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, H, W, n_pix, n_chunks);
    else
        tfy_launch_pdl((tfy_conv3x3_c1_fwd_tc_kernel<__nv_bfloat16>), dim3(grid), dim3(F1_THREADS), F1_SMEM, s,
                       (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y,
                       H, W, n_pix, n_chunks);
    return (int)cudaGetLastError();
}

// First-layer (C_in = 1, 32 filters) weight + bias gradient.  x: [B, H, W, 1] fp32 or bf16, dz: [B, H-2, W-2, 32] bf16
// (already gated by the layer's ReLU), dw: [32, 3, 3, 1] bf16, dbias: [32] bf16.  acc: 16 x 320 floats and counter:
// one uint32, both zero on entry (and left zero).
int tfy_conv3x3_c1_wgrad_tc(const void* x, int x_is_f32, const void* dz, float* acc, uint32_t* counter, void* dw,
                            void* dbias, int B, int H, int W, cudaStream_t s) {
    if (!c_init()) return -4;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tfy_conv3x3_c1_wgrad_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C1_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(tfy_conv3x3_c1_wgrad_tc_kernel<__nv_bfloat16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C1_SMEM) != cudaSuccess)
            return -5;
        attr = true;
    }
    const long long n_pix_ll = (long long)B * (H - 2) * (W - 2);
    if (n_pix_ll >= (1ll << 31) - C1_PIX) return -3;
    const int n_pix = (int)n_pix_ll, n_chunks = (n_pix + C1_PIX - 1) / C1_PIX;
    CUtensorMap mz;
    {
        cuuint64_t dims[2] = {(cuuint64_t)C1_O, (cuuint64_t)n_pix};
        cuuint64_t strides[1] = {(cuuint64_t)C1_O * 2};
        cuuint32_t box[2] = {(cuuint32_t)C1_O, (cuuint32_t)C1_PIX};
        cuuint32_t estr[2] = {1, 1};
        if (c_encode(&mz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dz), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -6;
    }
    const int grid = c_grid(n_chunks);
    if (x_is_f32)
        tfy_launch_pdl((tfy_conv3x3_c1_wgrad_tc_kernel<float>), dim3(grid), dim3(C1_THREADS), C1_SMEM, s, 
            mz, (const float*)x, acc, counter, (__nv_bfloat16*)dw, (__nv_bfloat16*)dbias, H, W, n_pix, n_chunks);
    else
        tfy_launch_pdl((tfy_conv3x3_c1_wgrad_tc_kernel<__nv_bfloat16>), dim3(grid), dim3(C1_THREADS), C1_SMEM, s, 
            mz, (const __nv_bfloat16*)x, acc, counter, (__nv_bfloat16*)dw, (__nv_bfloat16*)dbias, H, W, n_pix,
            n_chunks);
    return (int)cudaGetLastError();
}

}  // extern "C"
