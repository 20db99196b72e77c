// Parameter-server data plane over NVLink peer memory (sm_100a): K5 pull / K6 push.
//
// A ps rank owns shards in ITS HBM (inside the symmetric arena); chief/worker ranks map them and
// run these kernels against the peer addresses -- there is no server thread on the data path
// (reference path being replaced: TF gRPC RecvTensor pulls and Apply*/SparseApply*/ScatterAdd on the
// ps task, tf_yarn/tensorflow/cluster.py:41-67, tf_task_common.py:46-50).
//
//   pull   dense : peer fp32 master -> local replica (fp32 or bf16), 16-byte peer loads
//          sparse: gather the rows of a batch from a peer embedding table, fused with the bag
//                  reduction (sum / mean) that consumes them
//          GEMM  : weights stay remote and are streamed by TMA into the tcgen05 GEMM (tfy_gemm.cu)
//   push   dense : every worker applies its gradient to the peer master with red/atom -- asynchronous,
//                  lock free (TF's use_locking=False semantics), optimizer fused:
//                    SGD      w   += -lr*g                         (red.v4.f32)
//                    Adagrad  acc += g^2 ; w += -lr*g/(sqrt(acc)+eps)   (atom.v4.f32 returns old acc)
//                  and refreshes the bf16 shadow copy that the GEMM pull reads
//          sparse: the same per touched embedding row (duplicate ids handled by the atomics)
#include "tfy_common.cuh"

struct TfyPsSeg {
    uint64_t remote_w;       // peer fp32 master
    uint64_t remote_s1;      // peer fp32 optimizer slot (Adagrad accumulator), 0 if none
    uint64_t remote_shadow;  // peer bf16 shadow of the master, 0 if none
    uint64_t local;          // local replica (pull destination) / local gradient (push source)
    uint64_t n;              // elements
};

namespace {

__device__ __forceinline__ float4 ld_peer_f32x4(const void* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void red_add_f32x4(void* p, float4 v) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 atom_add_f32x4(void* p, float4 v) {
    float4 o;
    asm volatile("atom.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4], {%5,%6,%7,%8};"
                 : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                 : "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
    return o;
}
__device__ __forceinline__ void st_shadow_bf16x4(void* p, float4 v) {
    uint2 u;
    u.x = tfy_pack_bf16x2(v.x, v.y);
    u.y = tfy_pack_bf16x2(v.z, v.w);
    asm volatile("st.global.relaxed.sys.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(u.x), "r"(u.y) : "memory");
}

template <typename T>
__device__ __forceinline__ float4 ld_local4(const T* p);
template <>
__device__ __forceinline__ float4 ld_local4<float>(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ld_local4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(tfy_bf16lo(u.x), tfy_bf16hi(u.x), tfy_bf16lo(u.y), tfy_bf16hi(u.y));
}
template <typename T>
__device__ __forceinline__ void st_local4(T* p, float4 v);
template <>
__device__ __forceinline__ void st_local4<float>(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void st_local4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = tfy_pack_bf16x2(v.x, v.y);
    u.y = tfy_pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

}  // namespace

// ------------------------------------------------------------------------------------- dense pull
template <typename LT>
__global__ void __launch_bounds__(256) tfy_ps_pull_kernel(const TfyPsSeg* __restrict__ segs) {
    const TfyPsSeg sg = segs[blockIdx.y];
    const float* src = reinterpret_cast<const float*>(sg.remote_w);
    LT* dst = reinterpret_cast<LT*>(sg.local);
    const size_t n4 = sg.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        st_local4<LT>(dst + i * 4, ld_peer_f32x4(src + i * 4));
    if (blockIdx.x == 0)   // scalar tail (n % 4 elements)
        for (size_t i = n4 * 4 + threadIdx.x; i < sg.n; i += blockDim.x) dst[i] = (LT)src[i];
}

// ------------------------------------------------------------------------------------- dense push
// opt: 0 SGD, 3 Adagrad (TfyOpt numbering).  grad_scale multiplies the local gradient first.
template <typename GT, int OPT>
__global__ void __launch_bounds__(256)
tfy_ps_push_kernel(const TfyPsSeg* __restrict__ segs, float lr, float eps, float wd, float grad_scale) {
    const TfyPsSeg sg = segs[blockIdx.y];
    float* w = reinterpret_cast<float*>(sg.remote_w);
    float* acc = reinterpret_cast<float*>(sg.remote_s1);
    __nv_bfloat16* shadow = reinterpret_cast<__nv_bfloat16*>(sg.remote_shadow);
    const GT* g = reinterpret_cast<const GT*>(sg.local);
    const size_t n4 = sg.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 gv = ld_local4<GT>(g + i * 4);
        gv.x *= grad_scale; gv.y *= grad_scale; gv.z *= grad_scale; gv.w *= grad_scale;
        if (wd != 0.f) {
            const float4 wv = ld_peer_f32x4(w + i * 4);
            gv.x += wd * wv.x; gv.y += wd * wv.y; gv.z += wd * wv.z; gv.w += wd * wv.w;
        }
        float4 delta;
        if (OPT == TFY_OPT_ADAGRAD) {
            const float4 g2 = make_float4(gv.x * gv.x, gv.y * gv.y, gv.z * gv.z, gv.w * gv.w);
            const float4 old = atom_add_f32x4(acc + i * 4, g2);
            delta.x = -lr * gv.x / (sqrtf(old.x + g2.x) + eps);
            delta.y = -lr * gv.y / (sqrtf(old.y + g2.y) + eps);
            delta.z = -lr * gv.z / (sqrtf(old.z + g2.z) + eps);
            delta.w = -lr * gv.w / (sqrtf(old.w + g2.w) + eps);
        } else {
            delta = make_float4(-lr * gv.x, -lr * gv.y, -lr * gv.z, -lr * gv.w);
        }
        if (shadow) {
            const float4 old = atom_add_f32x4(w + i * 4, delta);
            st_shadow_bf16x4(shadow + i * 4,
                             make_float4(old.x + delta.x, old.y + delta.y, old.z + delta.z, old.w + delta.w));
        } else {
            red_add_f32x4(w + i * 4, delta);
        }
    }
    if (blockIdx.x == 0) {   // scalar tail (n % 4 elements)
        for (size_t i = n4 * 4 + threadIdx.x; i < sg.n; i += blockDim.x) {
            float gv = (float)g[i] * grad_scale;
            if (wd != 0.f) gv += wd * w[i];
            float delta;
            if (OPT == TFY_OPT_ADAGRAD) {
                const float old = atomicAdd(acc + i, gv * gv);
                delta = -lr * gv / (sqrtf(old + gv * gv) + eps);
            } else {
                delta = -lr * gv;
            }
            const float oldw = atomicAdd(w + i, delta);
            if (shadow) shadow[i] = __float2bfloat16(oldw + delta);
        }
    }
}

// ------------------------------------------------------------------------------------ sparse pull
// out[b, :] = reduce_j table[ids[b, j], :]   (mode 0 sum, 1 mean); table is a PEER fp32 [V, D], D % 4 == 0.
template <typename OT>
__global__ void __launch_bounds__(256)
tfy_ps_embedding_bag_kernel(const float* __restrict__ table, const long long* __restrict__ ids, OT* __restrict__ out,
                            int B, int L, int D, long long V, int mean) {
    const int D4 = D / 4;
    const size_t total = (size_t)B * D4;
    const float scale = mean ? 1.f / (float)L : 1.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / D4, d = (i % D4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < L; ++j) {
            long long id = ids[(size_t)b * L + j];
            if (id < 0 || id >= V) continue;
            const float4 v = ld_peer_f32x4(table + (size_t)id * D + d);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        st_local4<OT>(out + (size_t)b * D + d, acc);
    }
}

// ------------------------------------------------------------------------------------ sparse push
// For every (b, j): row = ids[b, j]; g = dout[b, :] * (mean ? 1/L : 1) applied to the peer row.
template <typename GT, int OPT>
__global__ void __launch_bounds__(256)
tfy_ps_push_rows_kernel(float* __restrict__ table, float* __restrict__ acc_table, const long long* __restrict__ ids,
                        const GT* __restrict__ dout, int B, int L, int D, long long V, int mean, float lr, float eps,
                        float grad_scale) {
    const int D4 = D / 4;
    const size_t total = (size_t)B * L * D4;
    const float scale = (mean ? 1.f / (float)L : 1.f) * grad_scale;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (i % D4) * 4;
        const size_t bj = i / D4;
        const int b = bj / L;
        const long long id = ids[bj];
        if (id < 0 || id >= V) continue;
        float4 gv = ld_local4<GT>(dout + (size_t)b * D + d);
        gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
        float* wrow = table + (size_t)id * D + d;
        float4 delta;
        if (OPT == TFY_OPT_ADAGRAD) {
            const float4 g2 = make_float4(gv.x * gv.x, gv.y * gv.y, gv.z * gv.z, gv.w * gv.w);
            const float4 old = atom_add_f32x4(acc_table + (size_t)id * D + d, g2);
            delta.x = -lr * gv.x / (sqrtf(old.x + g2.x) + eps);
            delta.y = -lr * gv.y / (sqrtf(old.y + g2.y) + eps);
            delta.z = -lr * gv.z / (sqrtf(old.z + g2.z) + eps);
            delta.w = -lr * gv.w / (sqrtf(old.w + g2.w) + eps);
        } else {
            delta = make_float4(-lr * gv.x, -lr * gv.y, -lr * gv.z, -lr * gv.w);
        }
        red_add_f32x4(wrow, delta);
    }
}

// fp32 -> bf16 shadow refresh of a whole segment (run by the chief after initialisation / restore)
__global__ void __launch_bounds__(256) tfy_ps_refresh_shadow_kernel(const TfyPsSeg* __restrict__ segs) {
    const TfyPsSeg sg = segs[blockIdx.y];
    if (!sg.remote_shadow) return;
    const float* w = reinterpret_cast<const float*>(sg.remote_w);
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(sg.remote_shadow);
    const size_t n4 = sg.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        st_shadow_bf16x4(sh + i * 4, ld_peer_f32x4(w + i * 4));
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < sg.n; i += blockDim.x) sh[i] = __float2bfloat16(w[i]);
}

extern "C" {

int tfy_ps_pull(const TfyPsSeg* segs, int nseg, size_t max_n, int local_is_bf16, cudaStream_t s) {
    if (nseg <= 0) return 0;
    size_t gx = (max_n / 4 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid((unsigned)gx, nseg);
    if (local_is_bf16) tfy_ps_pull_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(segs);
    else tfy_ps_pull_kernel<float><<<grid, 256, 0, s>>>(segs);
    return (int)cudaGetLastError();
}

int tfy_ps_push(const TfyPsSeg* segs, int nseg, size_t max_n, int grad_is_bf16, int opt, float lr, float eps, float wd,
                float grad_scale, cudaStream_t s) {
    if (nseg <= 0) return 0;
    if (opt != TFY_OPT_SGD && opt != TFY_OPT_ADAGRAD) return -2;
    size_t gx = (max_n / 4 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid((unsigned)gx, nseg);
#define TFY_PUSH(GT, O) tfy_ps_push_kernel<GT, O><<<grid, 256, 0, s>>>(segs, lr, eps, wd, grad_scale)
    if (grad_is_bf16) {
        if (opt == TFY_OPT_SGD) TFY_PUSH(__nv_bfloat16, TFY_OPT_SGD);
        else TFY_PUSH(__nv_bfloat16, TFY_OPT_ADAGRAD);
    } else {
        if (opt == TFY_OPT_SGD) TFY_PUSH(float, TFY_OPT_SGD);
        else TFY_PUSH(float, TFY_OPT_ADAGRAD);
    }
#undef TFY_PUSH
    return (int)cudaGetLastError();
}

int tfy_ps_refresh_shadow(const TfyPsSeg* segs, int nseg, size_t max_n, cudaStream_t s) {
    if (nseg <= 0) return 0;
    size_t gx = (max_n / 4 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid((unsigned)gx, nseg);
    tfy_ps_refresh_shadow_kernel<<<grid, 256, 0, s>>>(segs);
    return (int)cudaGetLastError();
}

int tfy_ps_embedding_bag(const void* table, const void* ids, void* out, int out_is_bf16, int B, int L, int D,
                         long long V, int mean, cudaStream_t s) {
    if (D % 4) return -2;
    size_t gx = ((size_t)B * (D / 4) + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
    if (out_is_bf16)
        tfy_ps_embedding_bag_kernel<__nv_bfloat16><<<(unsigned)gx, 256, 0, s>>>(
            (const float*)table, (const long long*)ids, (__nv_bfloat16*)out, B, L, D, V, mean);
    else
        tfy_ps_embedding_bag_kernel<float><<<(unsigned)gx, 256, 0, s>>>((const float*)table, (const long long*)ids,
                                                                        (float*)out, B, L, D, V, mean);
    return (int)cudaGetLastError();
}

int tfy_ps_push_rows(void* table, void* acc_table, const void* ids, const void* dout, int grad_is_bf16, int B, int L,
                     int D, long long V, int mean, int opt, float lr, float eps, float grad_scale, cudaStream_t s) {
    if (D % 4) return -2;
    if (opt != TFY_OPT_SGD && opt != TFY_OPT_ADAGRAD) return -3;
    if (opt == TFY_OPT_ADAGRAD && !acc_table) return -4;
    size_t gx = ((size_t)B * L * (D / 4) + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
#define TFY_PR(GT, O)                                                                                                 \
    tfy_ps_push_rows_kernel<GT, O><<<(unsigned)gx, 256, 0, s>>>((float*)table, (float*)acc_table,                    \
                                                                (const long long*)ids, (const GT*)dout, B, L, D, V,  \
                                                                mean, lr, eps, grad_scale)
    if (grad_is_bf16) {
        if (opt == TFY_OPT_SGD) TFY_PR(__nv_bfloat16, TFY_OPT_SGD);
        else TFY_PR(__nv_bfloat16, TFY_OPT_ADAGRAD);
    } else {
        if (opt == TFY_OPT_SGD) TFY_PR(float, TFY_OPT_SGD);
        else TFY_PR(float, TFY_OPT_ADAGRAD);
    }
#undef TFY_PR
    return (int)cudaGetLastError();
}

}  // extern "C"
