// Fused layer kernels of the mini-Keras fast path (sm_100a, bf16 activations, NHWC).
//
// The stock-PyTorch rendition of the MNIST-CNN step launches 53 kernels, most of
// them tiny elementwise / reduction kernels around the cuDNN/cuBLAS calls (bias
// add, clamp, max-pool fwd/bwd, dropout, bias-gradient reductions, the softmax /
// NLL chain, gradient accumulation adds) -- see profiles/launches_r1a_*.csv.
// These kernels fuse each such chain into ONE launch and write parameter
// gradients straight into the flat symmetric gradient buffer that the K4
// reduce-scatter/optimizer/all-gather kernel consumes:
//
//   tfy_conv3x3_c1_fwd        direct 3x3 conv for C_in = 1 + bias + ReLU
//   tfy_conv3x3_c1_wgrad      its weight gradient (reduction over all output pixels)
//   tfy_bias_act_drop_fwd     y = dropout(act(z + bias)) over [rows, C], keep-mask byte
//   tfy_act_drop_bwd_bias     dz = dy * mask (+ dbias[C] reduced and stored as bf16)
//   tfy_bias_relu_pool_drop_fwd   z -> bias + ReLU + 2x2 max-pool + dropout (+1 code byte)
//   tfy_pool_drop_relu_bwd        its backward (+ dbias[C])
//   tfy_softmax_xent          logits(+bias) -> mean loss, dlogits = (softmax - onehot)/B, dbias, #correct
//
// Dropout randomness is counter based: hash(seed, layer salt, optimizer step, element) with the
// step read from the device-resident TfyOptHyper, so a replayed CUDA graph draws fresh masks.
#include "tfy_common.cuh"

namespace {

__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// Cross-CTA column sums without a serial tail: every CTA adds its C block sums (in shared memory)
// into one of TFY_ACC_GROUPS global fp32 accumulator rows with RED (spreading the CTAs over several rows
// keeps the same-address contention at the L2 atomic units low); the CTA that arrives last swaps every
// accumulator with zero (atomicExch) -- which reads the totals AND re-arms the buffer for the next
// launch / graph replay -- and stores the bf16 result.  `gacc` (>= TFY_ACC_GROUPS*C floats) and
// `counter` must be zero on entry.
#define TFY_ACC_GROUPS 16
__device__ void tfy_colsum_publish(const float* s_sum, int C, float* gacc, __nv_bfloat16* out, uint32_t* counter) {
    __shared__ bool is_last;
    float* row = gacc + (size_t)(blockIdx.x % TFY_ACC_GROUPS) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&row[c], s_sum[c]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < TFY_ACC_GROUPS; ++g) t += atomicExch(&gacc[(size_t)g * C + c], 0.f);
        out[c] = __float2bfloat16(t);
    }
    if (threadIdx.x == 0) *counter = 0;
}

// Per-CTA column sums of values held 8-per-thread (column group g of G = C/8): reduce the lanes of a
// warp that share a group with shuffles, combine the warps through shared memory (s_sum: [C] floats,
// zeroed by the caller before use), then publish this CTA's row of `partial`.
__device__ __forceinline__ void tfy_block_colsum(float (&colacc)[8], int g, int G, bool active, float* s_sum) {
    if (G < 32 && (32 % G) == 0 && (blockDim.x % 32) == 0) {
        // lanes l and l + k*G of a warp hold the same column group (the grid stride keeps g = tid % G)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = active ? colacc[k] : 0.f;
            for (int off = G; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            colacc[k] = v;
        }
        if ((threadIdx.x & 31) < G) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(&s_sum[g * 8 + k], colacc[k]);
        }
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_sum[g * 8 + k], colacc[k]);
    }
    __syncthreads();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// conv 3x3, C_in = 1, stride 1, VALID, + bias + ReLU.   x: [B,H,W] (fp32 or bf16), w: [O][9] bf16,
// y: [B,H-2,W-2,O] bf16.  One thread = one output pixel, all O channels: x loads are coalesced across
// the warp, the weights are warp-broadcast reads from shared memory, each thread writes O*2 contiguous
// bytes.
// ---------------------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(256)
tfy_conv3x3_c1_fwd_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                          const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                          int O) {
    tfy_pdl_sync();
    extern __shared__ float s_w[];  // [9][O] + [O] bias
    for (int i = threadIdx.x; i < 9 * O; i += blockDim.x) {
        const int o = i % O, t = i / O;
        s_w[i] = bf16_to_f(w[o * 9 + t]);
    }
    for (int i = threadIdx.x; i < O; i += blockDim.x) s_w[9 * O + i] = bf16_to_f(bias[i]);
    __syncthreads();
    const int OH = H - 2, OW = W - 2;
    const size_t total = (size_t)B * OH * OW;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
         pix += (size_t)gridDim.x * blockDim.x) {
        const int ow = pix % OW;
        const size_t r = pix / OW;
        const int oh = r % OH;
        const int b = r / OH;
        const XT* xp = x + ((size_t)b * H + oh) * W + ow;
        float xin[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) xin[kh * 3 + kw] = (float)xp[kh * W + kw];
        __nv_bfloat16* yp = y + pix * O;
        for (int g = 0; g < O; g += 8) {
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = s_w[9 * O + g + k];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(xin[t], s_w[t * O + g + k], acc[k]);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.f);
            tfy_st16(yp + g, TfyPack<__nv_bfloat16>::pack(acc));
        }
    }
}

// dW[o][t] = sum over (b,oh,ow) x[b,oh+kh,ow+kw] * dz[b,oh,ow,o].
// lane <-> output channel (blocks of 32 channels).  A warp owns whole output rows (b, oh): it loads the
// three input rows it needs ONCE (coalesced, lane = column; W <= 32 fast path) and walks ow with the
// window values coming from warp shuffles -- per pixel one 64-byte dz load, 9 shuffles, 9 FMAs and no
// integer division.  Warps are combined in shared memory, CTAs through tfy_colsum_publish.
template <typename XT>
__global__ void __launch_bounds__(256)
tfy_conv3x3_c1_wgrad_kernel(const XT* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                            float* __restrict__ partial, __nv_bfloat16* __restrict__ dw, uint32_t* counter, int B,
                            int H, int W, int O) {
    extern __shared__ float s_acc[];   // [9*O] CTA accumulator
    const int OH = H - 2, OW = W - 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 9 * O; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int nrows = B * OH;
    const int gw = blockIdx.x * nwarps + warp, tw = gridDim.x * nwarps;
    for (int o0 = 0; o0 < O; o0 += 32) {
        const int o = o0 + lane;
        const bool o_ok = o < O;
        float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int row = gw; row < nrows; row += tw) {
            const int b = row / OH, oh = row - b * OH;
            const XT* xr = x + ((size_t)b * H + oh) * W;
            const __nv_bfloat16* dzr = dz + (size_t)row * OW * O + o;
            if (W <= 32) {
                const float r0 = lane < W ? (float)xr[lane] : 0.f;
                const float r1 = lane < W ? (float)xr[W + lane] : 0.f;
                const float r2 = lane < W ? (float)xr[2 * W + lane] : 0.f;
#pragma unroll 2
                for (int ow = 0; ow < OW; ++ow) {
                    const float d = o_ok ? bf16_to_f(dzr[(size_t)ow * O]) : 0.f;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        acc[0 + kw] = fmaf(__shfl_sync(0xffffffffu, r0, ow + kw), d, acc[0 + kw]);
                        acc[3 + kw] = fmaf(__shfl_sync(0xffffffffu, r1, ow + kw), d, acc[3 + kw]);
                        acc[6 + kw] = fmaf(__shfl_sync(0xffffffffu, r2, ow + kw), d, acc[6 + kw]);
                    }
                }
            } else {
                for (int ow = 0; ow < OW; ++ow) {
                    const float d = o_ok ? bf16_to_f(dzr[(size_t)ow * O]) : 0.f;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            acc[kh * 3 + kw] = fmaf((float)xr[kh * W + ow + kw], d, acc[kh * 3 + kw]);
                }
            }
        }
        if (o_ok) {
#pragma unroll
            for (int t = 0; t < 9; ++t) atomicAdd(&s_acc[o * 9 + t], acc[t]);
        }
    }
    __syncthreads();
    tfy_colsum_publish(s_acc, 9 * O, partial, dw, counter);
}

// ---------------------------------------------------------------------------------------------
// y = dropout(act(z + bias[c])) over a [rows, C] bf16 matrix (C % 8 == 0); y may alias z.
// mask (optional, 1 byte per element): 1 = gradient flows (act'(.) != 0 and kept).
// ---------------------------------------------------------------------------------------------
// ZT = __nv_bfloat16: z is a bf16 matrix.  ZT = float: z is the fp32 split-K accumulator of the tcgen05
// GEMM; it is consumed AND cleared here so the next step's partial sums start from zero.
template <typename ZT>
__global__ void __launch_bounds__(256)
tfy_bias_act_drop_fwd_kernel(ZT* z, const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* y,
                             uint8_t* __restrict__ mask, size_t rows, int C, int relu, float drop_rate, uint32_t seed,
                             const TfyOptHyper* __restrict__ hp) {
    tfy_pdl_sync();
    const int G = C / 8;
    const size_t total = rows * G;
    const uint32_t step = hp ? (uint32_t)hp->step : 0u;
    const float keep_scale = drop_rate > 0.f ? 1.f / (1.f - drop_rate) : 1.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % G;
        float v[8], bv[8];
        if (sizeof(ZT) == 4) {
            float4* zp = reinterpret_cast<float4*>(z) + idx * 2;
            const float4 a = zp[0], b = zp[1];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            zp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            zp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            TfyPack<__nv_bfloat16>::unpack(tfy_ld16(reinterpret_cast<const __nv_bfloat16*>(z) + idx * 8), v);
        }
        if (bias) {
            TfyPack<__nv_bfloat16>::unpack(tfy_ld16(bias + g * 8), bv);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += bv[k];
        }
        uint32_t m_lo = 0, m_hi = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bool on = true;
            if (relu) { on = v[k] > 0.f; v[k] = on ? v[k] : 0.f; }
            if (drop_rate > 0.f) {
                const bool keep = tfy_uniform(seed, step, idx * 8 + k) >= drop_rate;
                v[k] = keep ? v[k] * keep_scale : 0.f;
                on = on && keep;
            }
            if (k < 4) m_lo |= (on ? 1u : 0u) << (8 * k);
            else m_hi |= (on ? 1u : 0u) << (8 * (k - 4));
        }
        tfy_st16(y + idx * 8, TfyPack<__nv_bfloat16>::pack(v));
        if (mask) *reinterpret_cast<uint2*>(mask + idx * 8) = make_uint2(m_lo, m_hi);
    }
}

// dz = dy * gate * scale, where gate comes from `mask` bytes (if given) or from y > 0; dbias[c] = sum_rows dz.
__global__ void __launch_bounds__(256)
tfy_act_drop_bwd_bias_kernel(const __nv_bfloat16* dy, const uint8_t* __restrict__ mask,
                             const __nv_bfloat16* __restrict__ y, __nv_bfloat16* dz, float scale, size_t rows, int C,
                             float* __restrict__ partial, __nv_bfloat16* __restrict__ dbias, uint32_t* counter) {
    tfy_pdl_sync();
    extern __shared__ float s_sum[];  // [C]
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_sum[c] = 0.f;
    __syncthreads();
    const int G = C / 8;
    const size_t total = rows * G;
    // each thread keeps a fixed column group: stride must be a multiple of G
    const size_t stride = ((size_t)gridDim.x * blockDim.x / G) * G;
    float colacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = start < stride;
    const int g = (int)(start % G);
    if (active) {
        for (size_t idx = start; idx < total; idx += stride) {
            float v[8];
            TfyPack<__nv_bfloat16>::unpack(tfy_ld16(dy + idx * 8), v);
            if (mask) {
                const uint2 m = *reinterpret_cast<const uint2*>(mask + idx * 8);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t bit = (k < 4 ? (m.x >> (8 * k)) : (m.y >> (8 * (k - 4)))) & 1u;
                    v[k] = bit ? v[k] * scale : 0.f;
                }
            } else if (y) {
                float yv[8];
                TfyPack<__nv_bfloat16>::unpack(tfy_ld16(y + idx * 8), yv);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = yv[k] > 0.f ? v[k] * scale : 0.f;
            }
            if (dz) tfy_st16(dz + idx * 8, TfyPack<__nv_bfloat16>::pack(v));
#pragma unroll
            for (int k = 0; k < 8; ++k) colacc[k] += v[k];
        }
    }
    if (dbias) {
        tfy_block_colsum(colacc, g, G, active, s_sum);
        tfy_colsum_publish(s_sum, C, partial, dbias, counter);
    }
}

// ---------------------------------------------------------------------------------------------
// z [B,H,W,C] (conv output, pre-bias) -> p [B,H/2,W/2,C] = dropout(maxpool2x2(relu(z + bias)))
// code byte per pooled element: bits 0-1 = argmax position (dy*2+dx), bit 2 = gradient flows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tfy_bias_relu_pool_drop_fwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ bias,
                                   __nv_bfloat16* __restrict__ p, uint8_t* __restrict__ code, int B, int H, int W,
                                   int C, float drop_rate, uint32_t seed, const TfyOptHyper* __restrict__ hp) {
    const int PH = H / 2, PW = W / 2, G = C / 8;
    const size_t total = (size_t)B * PH * PW * G;
    const uint32_t step = hp ? (uint32_t)hp->step : 0u;
    const float keep_scale = drop_rate > 0.f ? 1.f / (1.f - drop_rate) : 1.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int g = idx % G;
        size_t r = idx / G;
        const int pw = r % PW;
        r /= PW;
        const int ph = r % PH;
        const int b = r / PH;
        float bv[8], best[8];
        int arg[8];
        TfyPack<__nv_bfloat16>::unpack(tfy_ld16(bias + g * 8), bv);
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -3.0e38f; arg[k] = 0; }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float v[8];
                const size_t off = (((size_t)b * H + ph * 2 + dy) * W + pw * 2 + dx) * C + g * 8;
                TfyPack<__nv_bfloat16>::unpack(tfy_ld16(z + off), v);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (v[k] > best[k]) { best[k] = v[k]; arg[k] = dy * 2 + dx; }
            }
        uint32_t c_lo = 0, c_hi = 0;
        float out[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = best[k] + bv[k];
            bool on = v > 0.f;
            v = on ? v : 0.f;
            if (drop_rate > 0.f) {
                const bool keep = tfy_uniform(seed, step, idx * 8 + k) >= drop_rate;
                v = keep ? v * keep_scale : 0.f;
                on = on && keep;
            }
            out[k] = v;
            const uint32_t cb = (uint32_t)arg[k] | (on ? 4u : 0u);
            if (k < 4) c_lo |= cb << (8 * k);
            else c_hi |= cb << (8 * (k - 4));
        }
        tfy_st16(p + idx * 8, TfyPack<__nv_bfloat16>::pack(out));
        *reinterpret_cast<uint2*>(code + idx * 8) = make_uint2(c_lo, c_hi);
    }
}

// dp [B,PH,PW,C] + code -> dz [B,H,W,C] (gradient wrt the conv output), dbias[C] = sum dz.
__global__ void __launch_bounds__(256)
tfy_pool_drop_relu_bwd_kernel(const __nv_bfloat16* __restrict__ dp, const uint8_t* __restrict__ code,
                              __nv_bfloat16* __restrict__ dz, float scale, int B, int H, int W, int C,
                              float* __restrict__ partial, __nv_bfloat16* __restrict__ dbias, uint32_t* counter) {
    tfy_pdl_sync();
    extern __shared__ float s_sum[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_sum[c] = 0.f;
    __syncthreads();
    const int PH = H / 2, PW = W / 2, G = C / 8;
    const size_t total = (size_t)B * PH * PW * G;
    const size_t stride = ((size_t)gridDim.x * blockDim.x / G) * G;
    const size_t start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = start < stride;
    const int g = (int)(start % G);
    float colacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
        for (size_t idx = start; idx < total; idx += stride) {
            size_t r = idx / G;
            const int pw = r % PW;
            r /= PW;
            const int ph = r % PH;
            const int b = r / PH;
            float d[8];
            TfyPack<__nv_bfloat16>::unpack(tfy_ld16(dp + idx * 8), d);
            const uint2 cd = *reinterpret_cast<const uint2*>(code + idx * 8);
            int arg[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t cb = (k < 4 ? (cd.x >> (8 * k)) : (cd.y >> (8 * (k - 4)))) & 0xffu;
                d[k] = (cb & 4u) ? d[k] * scale : 0.f;
                arg[k] = cb & 3u;
                colacc[k] += d[k];
            }
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = arg[k] == pos ? d[k] : 0.f;
                const size_t off = (((size_t)b * H + ph * 2 + (pos >> 1)) * W + pw * 2 + (pos & 1)) * C + g * 8;
                tfy_st16(dz + off, TfyPack<__nv_bfloat16>::pack(o));
            }
        }
    }
    if (dbias) {
        tfy_block_colsum(colacc, g, G, active, s_sum);
        tfy_colsum_publish(s_sum, C, partial, dbias, counter);
    }
}

// ---------------------------------------------------------------------------------------------
// softmax cross-entropy head.  logits [B,C] bf16 (+ optional bias[C]), labels int64.
// Single CTA (B rows strided over the threads): loss = mean CE (fp32), dlogits = (softmax-onehot)/B
// (bf16), dbias[C] = column sums of dlogits (bf16), stats[0] += #correct, stats[1] += B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
tfy_softmax_xent_kernel(const __nv_bfloat16* __restrict__ logits, const __nv_bfloat16* __restrict__ bias,
                        const long long* __restrict__ labels, float* __restrict__ loss,
                        __nv_bfloat16* __restrict__ dlogits, __nv_bfloat16* __restrict__ dbias,
                        float* __restrict__ stats, int B, int C) {
    extern __shared__ float sm[];  // [C] bias | [C] dbias accumulators | [2] loss, correct
    float* s_bias = sm;
    float* s_db = sm + C;
    float* s_red = sm + 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_bias[c] = bias ? bf16_to_f(bias[c]) : 0.f;
        s_db[c] = 0.f;
    }
    if (threadIdx.x < 2) s_red[threadIdx.x] = 0.f;
    __syncthreads();
    const float invB = 1.f / (float)B;
    const int lane = threadIdx.x & 31;
    float my_loss = 0.f, my_correct = 0.f;
    // every warp iterates the same number of times so that the shuffles below stay convergent
    const int rounds = (B + blockDim.x - 1) / blockDim.x;
    for (int it = 0; it < rounds; ++it) {
        const int r = it * blockDim.x + threadIdx.x;
        const bool ok = r < B;
        const __nv_bfloat16* row = logits + (size_t)(ok ? r : 0) * C;
        float mx = -3.0e38f;
        int amax = 0;
        for (int c = 0; c < C; ++c) {
            const float v = bf16_to_f(row[c]) + s_bias[c];
            if (v > mx) { mx = v; amax = c; }
        }
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += __expf(bf16_to_f(row[c]) + s_bias[c] - mx);
        const float lse = mx + __logf(se);
        const int lab = ok ? (int)labels[r] : 0;
        const float inv_se = 1.f / se;
        for (int c = 0; c < C; ++c) {
            const float v = bf16_to_f(row[c]) + s_bias[c];
            float d = (__expf(v - mx) * inv_se - (c == lab ? 1.f : 0.f)) * invB;
            if (!ok) d = 0.f;
            if (ok) {
                if (c == lab) my_loss += lse - v;
                dlogits[(size_t)r * C + c] = __float2bfloat16(d);
            }
            if (dbias) {
                float t = d;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane == 0) atomicAdd(&s_db[c], t);
            }
        }
        if (ok) my_correct += (amax == lab) ? 1.f : 0.f;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        my_loss += __shfl_xor_sync(0xffffffffu, my_loss, off);
        my_correct += __shfl_xor_sync(0xffffffffu, my_correct, off);
    }
    if (lane == 0) {
        atomicAdd(&s_red[0], my_loss);
        atomicAdd(&s_red[1], my_correct);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *loss = s_red[0] * invB;
        if (stats) { stats[0] += s_red[1]; stats[1] += (float)B; }
    }
    if (dbias)
        for (int c = threadIdx.x; c < C; c += blockDim.x) dbias[c] = __float2bfloat16(s_db[c]);
}


// ---------------------------------------------------------------------------------------------
// Classifier head, forward AND backward, in one launch (the whole problem is ~0.5 MFLOP):
//   logits = h W2^T + b2 ; loss = mean CE(softmax(logits), labels) ; dlogits = (p - onehot) / B
//   dW2 = dlogits^T h ; db2 = colsum(dlogits)
//   dh  = (dlogits W2) * gate(mask1) * scale1 ; db1 = colsum(dh)          (the previous Dense layer's
//                                                                          ReLU / dropout gate + bias gradient)
// Replaces a cuBLAS GEMM, the softmax/CE kernel, two more cuBLAS GEMMs and the gate/bias-gradient kernel
// (3.9 + 6.6 + 4.6 + 4.1 + 6.9 us of launches for 128 x 128 x 10).  Intermediates stay in fp32.
// One CTA per 16 rows; the batch reductions (dW2, db2, db1, loss) meet in an fp32 scratch buffer and the
// last CTA to finish converts and clears it.
// h: [B, K] bf16, W2: [C, K] bf16, b2: [C] bf16, labels: [B] int64, mask1: [B, K] uint8 (bit 0 = gradient
// flows) or null.  K in {128, 256, 512}, C <= 16.
// scratch: [C*K | 16 | K | 2] floats, zero on entry and on exit; counter: one uint32, same.
// ---------------------------------------------------------------------------------------------
// debug: per-CTA event timeline of the head kernel (16 slots of SM-clock deltas per CTA), see tfy_nn_set_timeline
__device__ long long* n_timeline = nullptr;
#define N_MARK(k)                                                                    \
    do {                                                                             \
        if (n_tl && threadIdx.x == 0) n_tl[(size_t)blockIdx.x * 16 + (k)] = clock64() - n_t0; \
    } while (0)

constexpr int HEAD_THREADS = 256, HEAD_CMAX = 16, HEAD_ROWS = 16;

template <int K>
__global__ void __launch_bounds__(HEAD_THREADS)
tfy_dense_head_fused_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ w2,
                            const __nv_bfloat16* __restrict__ b2, const long long* __restrict__ labels,
                            const uint8_t* __restrict__ mask1, float scale1, float* __restrict__ loss,
                            float* __restrict__ loss_host, float* __restrict__ stats,
                            __nv_bfloat16* __restrict__ dw2,
                            __nv_bfloat16* __restrict__ db2, __nv_bfloat16* __restrict__ dh,
                            __nv_bfloat16* __restrict__ db1, float* __restrict__ scratch,
                            uint32_t* __restrict__ counter, int B, int C) {
    // K is a template parameter: every loop that waits on global memory has a constant trip count and is
    // fully unrolled, so its loads are issued back to back (the first version had ~18 serialised L2 round
    // trips in five-iteration loops: 15 us for 0.5 MFLOP).
    tfy_pdl_sync();
    long long* const n_tl = n_timeline;
    const long long n_t0 = clock64();
    extern __shared__ __align__(16) uint8_t head_smem[];
    constexpr int KP = K + 8;                                       // padded row (bank spread)
    __nv_bfloat16* s_h = reinterpret_cast<__nv_bfloat16*>(head_smem);                 // [16][KP]
    float* s_w = reinterpret_cast<float*>(head_smem + (size_t)HEAD_ROWS * KP * 2);    // [16][K], rows >= C are zero
    float* s_dl = s_w + (size_t)HEAD_CMAX * K;                                                // [16][HEAD_CMAX]
    float* s_b = s_dl + HEAD_ROWS * HEAD_CMAX;                                        // [HEAD_CMAX]
    float* s_red = s_b + HEAD_CMAX;                                                   // loss, correct
    float* s_col = s_red + 4;                                                         // [max(K, 256)], 16-byte aligned
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * HEAD_ROWS;
    const int rows = min(HEAD_ROWS, B - row0);
    float* g_dw = scratch;
    float* g_db2 = scratch + (size_t)C * K;
    float* g_db1 = g_db2 + HEAD_CMAX;
    float* g_red = g_db1 + K;

    // ---- A: stage this CTA's rows of h (bf16), W2 and b2 (fp32)
    const int kv = K / 8;
    for (int i = tid; i < HEAD_ROWS * kv; i += HEAD_THREADS) {
        const int r = i / kv, g = i - r * kv;
        *reinterpret_cast<uint4*>(s_h + (size_t)r * KP + g * 8) =
            r < rows ? tfy_ld16(h + (size_t)(row0 + r) * K + g * 8) : make_uint4(0, 0, 0, 0);
    }
    {
        constexpr int WV = HEAD_CMAX * K / 8 / HEAD_THREADS;        // 16-byte packs of W2 per thread (upper bound)
        uint4 wv[WV];
#pragma unroll
        for (int q = 0; q < WV; ++q) {
            const int i = tid + q * HEAD_THREADS;
            wv[q] = i < C * K / 8 ? tfy_ld16(w2 + (size_t)i * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < WV; ++q) {                          // packs beyond C*K/8 were loaded as zeros
            const int i = tid + q * HEAD_THREADS;
            TfyPack<__nv_bfloat16>::unpack(wv[q], s_w + (size_t)i * 8);
        }
    }
    if (tid < HEAD_CMAX) s_b[tid] = (tid < C && b2) ? bf16_to_f(b2[tid]) : 0.f;
    if (tid < 2) s_red[tid] = 0.f;
    for (int i = tid; i < HEAD_ROWS * HEAD_CMAX; i += HEAD_THREADS) s_dl[i] = 0.f;
    __syncthreads();

    N_MARK(1);
    // ---- B/C: logits (16 threads per row), softmax, loss, dlogits
    const float invB = 1.f / (float)B;
    {
        // Classes are padded to 16 with zero weights (no per-class branches) and taken 4 at a time in a ROLLED
        // loop: this code runs once per launch on 8 warps, so it is instruction-fetch bound -- the fully
        // unrolled version (2400 instructions) spent 7k cycles here, mostly waiting for the I-cache.
        const int r = tid >> 4, part = tid & 15;
        constexpr int kq = K / 16;
        const int lab = r < rows ? (int)labels[row0 + r] : -1;     // issued early: used after the GEMV
        const __nv_bfloat16* hr = s_h + (size_t)r * KP + part * kq;
        float* s_logit = s_col;                                     // [16][16] scratch (s_col is used later)
#pragma unroll 1
        for (int cb = 0; cb < HEAD_CMAX; cb += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int k = 0; k < kq; k += 8) {
                float hv[8];
                TfyPack<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(hr + k), hv);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* wp = s_w + (size_t)(cb + c) * K + part * kq + k;
                    const float4 w0 = *reinterpret_cast<const float4*>(wp);
                    const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
                    acc[c] = fmaf(hv[0], w0.x, acc[c]); acc[c] = fmaf(hv[1], w0.y, acc[c]);
                    acc[c] = fmaf(hv[2], w0.z, acc[c]); acc[c] = fmaf(hv[3], w0.w, acc[c]);
                    acc[c] = fmaf(hv[4], w1.x, acc[c]); acc[c] = fmaf(hv[5], w1.y, acc[c]);
                    acc[c] = fmaf(hv[6], w1.z, acc[c]); acc[c] = fmaf(hv[7], w1.w, acc[c]);
                }
            }
#pragma unroll
            for (int off = 1; off < 16; off <<= 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
            }
            if (part == 0) *reinterpret_cast<float4*>(s_logit + r * HEAD_CMAX + cb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncwarp();
        // softmax / cross-entropy with one lane per class (the 16 lanes of a row group): a handful of shuffle
        // steps instead of three serial 10-iteration loops on one lane per row
        {
            const int c = part;
            const bool live = c < C && r < rows;
            const float v = live ? s_logit[r * HEAD_CMAX + c] + s_b[c] : -3.0e38f;
            float mx = v;
            int amax = c;
#pragma unroll
            for (int off = 1; off < 16; off <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, mx, off);
                const int oa = __shfl_xor_sync(0xffffffffu, amax, off);
                if (ov > mx || (ov == mx && oa < amax)) { mx = ov; amax = oa; }
            }
            const float e = live ? __expf(v - mx) : 0.f;
            float se = e;
#pragma unroll
            for (int off = 1; off < 16; off <<= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
            const bool hit = live && c == lab;
            if (live) s_dl[r * HEAD_CMAX + c] = (e / se - (hit ? 1.f : 0.f)) * invB;
            float my_loss = hit ? (mx + __logf(se)) - v : 0.f;
            float my_correct = (hit && amax == lab) ? 1.f : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                my_loss += __shfl_xor_sync(0xffffffffu, my_loss, off);
                my_correct += __shfl_xor_sync(0xffffffffu, my_correct, off);
            }
            if ((tid & 31) == 0) {
                atomicAdd(&s_red[0], my_loss);
                atomicAdd(&s_red[1], my_correct);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < K; i += HEAD_THREADS) s_col[i] = 0.f;     // s_logit scratch -> column sums
    __syncthreads();
    if (tid == 0) {
        atomicAdd(g_red, s_red[0]);
        atomicAdd(g_red + 1, s_red[1]);
    }

    N_MARK(2);
    // ---- D: partial dW2[c][k] = sum_r dl[r][c] h[r][k]; db2[c] = sum_r dl[r][c]
    {
        // a thread keeps one column k of h (16 values) in registers and runs over the classes it owns
        constexpr int KREP = K >= HEAD_THREADS ? K / HEAD_THREADS : 1;
        constexpr int CSTEP = K >= HEAD_THREADS ? 1 : HEAD_THREADS / K;
#pragma unroll
        for (int kr = 0; kr < KREP; ++kr) {
            const int k = (K >= HEAD_THREADS) ? tid + kr * HEAD_THREADS : tid % K;
            float hc[HEAD_ROWS];
#pragma unroll
            for (int r = 0; r < HEAD_ROWS; ++r) hc[r] = bf16_to_f(s_h[(size_t)r * KP + k]);
#pragma unroll 1
            for (int cc = 0; cc < HEAD_CMAX / CSTEP; ++cc) {
                const int c = cc * CSTEP + ((K >= HEAD_THREADS) ? 0 : tid / K);
                if (c >= C) break;
                float a = 0.f;
#pragma unroll
                for (int r = 0; r < HEAD_ROWS; ++r) a = fmaf(s_dl[r * HEAD_CMAX + c], hc[r], a);
                if (c < C) atomicAdd(g_dw + (size_t)c * K + k, a);
            }
        }
    }
    if (tid < C) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < HEAD_ROWS; ++r) a += s_dl[r * HEAD_CMAX + tid];
        atomicAdd(g_db2 + tid, a);
    }

    N_MARK(3);
    // ---- E: dh[r][k] = (sum_c dl[r][c] W2[c][k]) * gate * scale; db1[k] = sum_r dh[r][k]
    if (dh) {
        // a thread keeps one column k of W2 (16 classes) in registers and runs over the rows it owns
        constexpr int KREP = K >= HEAD_THREADS ? K / HEAD_THREADS : 1;
        constexpr int RSTEP = K >= HEAD_THREADS ? 1 : HEAD_THREADS / K;
#pragma unroll
        for (int kr = 0; kr < KREP; ++kr) {
            const int k = (K >= HEAD_THREADS) ? tid + kr * HEAD_THREADS : tid % K;
            const int rbase = (K >= HEAD_THREADS) ? 0 : tid / K;
            float wk[HEAD_CMAX];
#pragma unroll
            for (int c = 0; c < HEAD_CMAX; ++c) wk[c] = s_w[(size_t)c * K + k];
            uint32_t mkbits = 0;                                    // gate bits of the rows this thread owns
#pragma unroll
            for (int j = 0; j < HEAD_ROWS / RSTEP; ++j) {
                const int r = rbase + j * RSTEP;
                const uint32_t bit = (mask1 && r < rows) ? (uint32_t)(mask1[(size_t)(row0 + r) * K + k] & 1) : 1u;
                mkbits |= bit << j;
            }
            float col = 0.f;
#pragma unroll 2
            for (int j = 0; j < HEAD_ROWS / RSTEP; ++j) {
                const int r = rbase + j * RSTEP;
                float v = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < HEAD_CMAX; c4 += 4) {
                    const float4 d = *reinterpret_cast<const float4*>(s_dl + r * HEAD_CMAX + c4);
                    v = fmaf(d.x, wk[c4], v); v = fmaf(d.y, wk[c4 + 1], v);
                    v = fmaf(d.z, wk[c4 + 2], v); v = fmaf(d.w, wk[c4 + 3], v);
                }
                if (mask1) v = ((mkbits >> j) & 1u) ? v * scale1 : 0.f;
                if (r < rows) {
                    dh[(size_t)(row0 + r) * K + k] = __float2bfloat16(v);
                    col += v;
                }
            }
            if (db1) atomicAdd(&s_col[k], col);
        }
        if (db1) {
            __syncthreads();
            for (int k = tid; k < K; k += HEAD_THREADS) atomicAdd(g_db1 + k, s_col[k]);
        }
    }

    N_MARK(4);
    // ---- the last CTA converts the batch sums and clears the scratch buffer
    __threadfence();
    __syncthreads();
    N_MARK(5);
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    N_MARK(6);
    if (s_last) {
        __threadfence();
        {
            constexpr int FV = HEAD_CMAX * K / 4 / HEAD_THREADS;    // float4 packs per thread (upper bound)
            float4 fv[FV];
#pragma unroll
            for (int q = 0; q < FV; ++q) {
                const int i = tid + q * HEAD_THREADS;
                fv[q] = i < C * K / 4 ? __ldcg(reinterpret_cast<const float4*>(g_dw) + i) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int q = 0; q < FV; ++q) {
                const int i = tid + q * HEAD_THREADS;
                if (i < C * K / 4) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(fv[q].x, fv[q].y), hi = __floats2bfloat162_rn(fv[q].z, fv[q].w);
                    uint2 packed;
                    packed.x = *reinterpret_cast<uint32_t*>(&lo);
                    packed.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(dw2 + (size_t)i * 4) = packed;
                    __stcg(reinterpret_cast<float4*>(g_dw) + i, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
        }
        if (tid < HEAD_CMAX) {
            if (tid < C && db2) db2[tid] = __float2bfloat16(__ldcg(g_db2 + tid));
            __stcg(g_db2 + tid, 0.f);
        }
        for (int k = tid; k < K; k += HEAD_THREADS) {
            if (db1) db1[k] = __float2bfloat16(__ldcg(g_db1 + k));
            __stcg(g_db1 + k, 0.f);
        }
        if (tid == 0) {
            const float mean_loss = __ldcg(g_red) * invB;
            *loss = mean_loss;
            if (loss_host) *loss_host = mean_loss;      // pinned host scalar (UVA): no D2H copy node per step
            if (stats) { stats[0] += __ldcg(g_red + 1); stats[1] += (float)B; }
            __stcg(g_red, 0.f);
            __stcg(g_red + 1, 0.f);
            *counter = 0u;
        }
    }
    N_MARK(7);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline int tfy_grid_for(size_t items, int block, int cap) {
    size_t g = (items + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

extern "C" {

// scratch: fp32 workspace of at least tfy_nn_scratch_floats(...) floats; counter: zero-initialised u32
size_t tfy_nn_max_partial_blocks() { return 592; }

int tfy_conv3x3_c1_fwd(const void* x, int x_is_f32, const void* w, const void* bias, void* y, int B, int H, int W,
                       int O, cudaStream_t s) {
    if (O % 8) return -2;
    const size_t total = (size_t)B * (H - 2) * (W - 2);
    const int grid = tfy_grid_for(total, 256, 148 * 8);
    const size_t smem = (size_t)(10 * O) * sizeof(float);
    if (x_is_f32)
        tfy_launch_pdl((tfy_conv3x3_c1_fwd_kernel<float>), dim3(grid), dim3(256), smem, s, (const float*)x, (const __nv_bfloat16*)w,
                                                                  (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, B, H,
                                                                  W, O);
    else
        tfy_launch_pdl((tfy_conv3x3_c1_fwd_kernel<__nv_bfloat16>), dim3(grid), dim3(256), smem, s, 
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, B, H, W, O);
    return (int)cudaGetLastError();
}

int tfy_conv3x3_c1_wgrad(const void* x, int x_is_f32, const void* dz, float* partial, void* dw, uint32_t* counter,
                         int B, int H, int W, int O, cudaStream_t s) {
    const int grid = 296;
    const size_t smem = (size_t)(9 * O > 256 ? 9 * O : 256) * sizeof(float);
    if (x_is_f32)
        tfy_conv3x3_c1_wgrad_kernel<float><<<grid, 256, smem, s>>>((const float*)x, (const __nv_bfloat16*)dz, partial,
                                                                    (__nv_bfloat16*)dw, counter, B, H, W, O);
    else
        tfy_conv3x3_c1_wgrad_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, partial, (__nv_bfloat16*)dw, counter, B, H, W, O);
    return (int)cudaGetLastError();
}

int tfy_bias_act_drop_fwd(const void* z, const void* bias, void* y, void* mask, size_t rows, int C, int relu,
                          float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    if (C % 8) return -2;
    const int grid = tfy_grid_for(rows * (C / 8), 256, 148 * 8);
    tfy_launch_pdl((tfy_bias_act_drop_fwd_kernel<__nv_bfloat16>), dim3(grid), dim3(256), 0, s, (__nv_bfloat16*)z, (const __nv_bfloat16*)bias,
                                                                     (__nv_bfloat16*)y, (uint8_t*)mask, rows, C, relu,
                                                                     drop_rate, seed, hp);
    return (int)cudaGetLastError();
}

// same, reading (and clearing) the fp32 split-K accumulator z32 [rows, C]
int tfy_bias_act_drop_fwd_f32(void* z32, const void* bias, void* y, void* mask, size_t rows, int C, int relu,
                              float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    if (C % 8) return -2;
    const int grid = tfy_grid_for(rows * (C / 8), 256, 148 * 8);
    tfy_launch_pdl((tfy_bias_act_drop_fwd_kernel<float>), dim3(grid), dim3(256), 0, s, (float*)z32, (const __nv_bfloat16*)bias,
                                                             (__nv_bfloat16*)y, (uint8_t*)mask, rows, C, relu,
                                                             drop_rate, seed, hp);
    return (int)cudaGetLastError();
}

int tfy_act_drop_bwd_bias(const void* dy, const void* mask, const void* y, void* dz, float scale, size_t rows, int C,
                          float* partial, void* dbias, uint32_t* counter, cudaStream_t s) {
    if (C % 8) return -2;
    int grid = tfy_grid_for(rows * (C / 8), 256, 592);
    // the kernel needs gridDim*blockDim >= C/8 so that every column group has a thread
    if ((size_t)grid * 256 < (size_t)(C / 8)) grid = (C / 8 + 255) / 256;
    tfy_launch_pdl((tfy_act_drop_bwd_bias_kernel), dim3(grid), dim3(256), (C > 256 ? C : 256) * sizeof(float), s, 
        (const __nv_bfloat16*)dy, (const uint8_t*)mask, (const __nv_bfloat16*)y, (__nv_bfloat16*)dz, scale, rows, C,
        partial, (__nv_bfloat16*)dbias, counter);
    return (int)cudaGetLastError();
}

int tfy_bias_relu_pool_drop_fwd(const void* z, const void* bias, void* p, void* code, int B, int H, int W, int C,
                                float drop_rate, uint32_t seed, const TfyOptHyper* hp, cudaStream_t s) {
    if (C % 8 || H % 2 || W % 2) return -2;
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    const int grid = tfy_grid_for(total, 256, 148 * 8);
    tfy_bias_relu_pool_drop_fwd_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)z, (const __nv_bfloat16*)bias,
                                                            (__nv_bfloat16*)p, (uint8_t*)code, B, H, W, C, drop_rate,
                                                            seed, hp);
    return (int)cudaGetLastError();
}

int tfy_pool_drop_relu_bwd(const void* dp, const void* code, void* dz, float scale, int B, int H, int W, int C,
                           float* partial, void* dbias, uint32_t* counter, cudaStream_t s) {
    if (C % 8 || H % 2 || W % 2) return -2;
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    int grid = tfy_grid_for(total, 256, 592);
    tfy_launch_pdl((tfy_pool_drop_relu_bwd_kernel), dim3(grid), dim3(256), (C > 256 ? C : 256) * sizeof(float), s, 
        (const __nv_bfloat16*)dp, (const uint8_t*)code, (__nv_bfloat16*)dz, scale, B, H, W, C, partial,
        (__nv_bfloat16*)dbias, counter);
    return (int)cudaGetLastError();
}

int tfy_softmax_xent(const void* logits, const void* bias, const void* labels, float* loss, void* dlogits,
                     void* dbias, float* stats, int B, int C, cudaStream_t s) {
    int block = B < 1024 ? ((B + 31) / 32) * 32 : 1024;
    if (block < 32) block = 32;
    tfy_softmax_xent_kernel<<<1, block, (2 * C + 2) * sizeof(float), s>>>(
        (const __nv_bfloat16*)logits, (const __nv_bfloat16*)bias, (const long long*)labels, loss,
        (__nv_bfloat16*)dlogits, (__nv_bfloat16*)dbias, stats, B, C);
    return (int)cudaGetLastError();
}

// thin cudaMemcpyAsync wrapper: the input pipeline issues its per-step copies through ctypes (about 1 us per call
// instead of ~10 us for Tensor.copy_ under a stream context manager)
int tfy_memcpy_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s);
}

int tfy_nn_set_timeline(long long* buf) { return (int)cudaMemcpyToSymbol(n_timeline, &buf, sizeof(buf)); }

// fused classifier head (forward + backward); returns -2 when the shape is outside the kernel's envelope.
// scratch: tfy_dense_head_scratch_elems(K, C) floats and counter: one uint32, zero on entry (left zero).
// loss_host: optional second destination of the loss, e.g. a pinned host scalar (device-accessible under UVA).
size_t tfy_dense_head_scratch_elems(int K, int C) { return (size_t)C * K + HEAD_CMAX + K + 2; }

int tfy_dense_head_fused(const void* h, const void* w2, const void* b2, const void* labels, const void* mask1,
                         float scale1, float* loss, float* loss_host, float* stats, void* dw2, void* db2, void* dh,
                         void* db1,
                         float* scratch, uint32_t* counter, int B, int K, int C, cudaStream_t s) {
    if (!(K == 128 || K == 256 || K == 512) || C < 1 || C > HEAD_CMAX || B < 1) return -2;
    const size_t smem = (size_t)HEAD_ROWS * (K + 8) * 2 +
                        ((size_t)HEAD_CMAX * K + HEAD_ROWS * HEAD_CMAX + HEAD_CMAX + 4 + (K > 256 ? K : 256)) * 4;
    const int grid = (B + HEAD_ROWS - 1) / HEAD_ROWS;
#define TFY_HEAD(KK)                                                                                                   \
    do {                                                                                                               \
        static bool configured = false;                                                                                \
        if (smem > 48 * 1024 && !configured) {                                                                         \
            if (cudaFuncSetAttribute(tfy_dense_head_fused_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     96 * 1024) != cudaSuccess)                                                        \
                return -5;                                                                                             \
            configured = true;                                                                                         \
        }                                                                                                              \
        tfy_launch_pdl((tfy_dense_head_fused_kernel<KK>), dim3(grid), dim3(HEAD_THREADS), smem, s,                     \
                       (const __nv_bfloat16*)h, (const __nv_bfloat16*)w2, (const __nv_bfloat16*)b2,                    \
                       (const long long*)labels, (const uint8_t*)mask1, scale1, loss, loss_host, stats,              \
                       (__nv_bfloat16*)dw2,                                                                            \
                       (__nv_bfloat16*)db2, (__nv_bfloat16*)dh, (__nv_bfloat16*)db1, scratch, counter, B, C);          \
    } while (0)
    if (K == 128) TFY_HEAD(128);
    else if (K == 256) TFY_HEAD(256);
    else TFY_HEAD(512);
#undef TFY_HEAD
    return (int)cudaGetLastError();
}

}  // extern "C"
