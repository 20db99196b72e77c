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
//                  lock free (TF's use_locking=False semantics), optimizer fused, chosen PER VARIABLE (wide-and-deep
//                  trains its linear tower with FTRL and its deep tower with Adagrad):
//                    SGD      w   += -lr*g                                       (red.v4.f32)
//                    Adagrad  acc += g^2 ; w += -lr*g/(sqrt(acc)+eps)            (atom.v4.f32 returns old acc)
//                    FTRL     n += g^2 ; z += g - (sqrt(n')-sqrt(n))/lr * w ;    (two atom.v4.f32)
//                             w  = |z|<=l1 ? 0 : -(z - sgn(z) l1) / ((beta+sqrt(n'))/lr + 2 l2)
//                    Adam     m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;      (hogwild read-modify-write of m, v)
//                             w += -lr_t m/(sqrt(v)+eps)                         (red.v4.f32)
//                  and refreshes the bf16 shadow copy that the GEMM pull reads
//          sparse: the same per touched embedding row (duplicate ids handled by the atomics)
#include "tfy_common.cuh"

struct TfyPsSeg {
    uint64_t remote_w;       // peer fp32 master
    uint64_t remote_s1;      // peer fp32 optimizer slot 1 (Adagrad / FTRL accumulator, Adam m), 0 if none
    uint64_t remote_s2;      // peer fp32 optimizer slot 2 (FTRL linear z, Adam v), 0 if none
    uint64_t remote_shadow;  // peer bf16 shadow of the master, 0 if none
    uint64_t local;          // local replica (pull destination) / local gradient (push source)
    uint64_t n;              // elements
    int32_t opt;             // TfyOpt of THIS variable
    float lr, eps, wd;
    float p1, p2, p3;        // FTRL: l1, l2, beta | Adam: beta1, beta2, -
    int32_t pad;
    // bf16 shadow of a 2-D weight [rows, cols]: rows are `shadow_ld` elements apart (cols rounded up to 8 so that
    // a TMA tensor map can describe it); cols == 0 means "flat, same indexing as the master"
    uint32_t cols, shadow_ld;
};

// optimizer hyper-parameters of one push (sparse rows: passed by value)
struct TfyPsHyper {
    int32_t opt;
    float lr, eps, wd, p1, p2, p3, grad_scale;
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
// element index of the master -> element index of the (row-padded) shadow
__device__ __forceinline__ size_t shadow_index(size_t i, uint32_t cols, uint32_t ld) {
    return cols ? (i / cols) * (size_t)ld + (i % cols) : i;
}
__device__ __forceinline__ void st_shadow4(__nv_bfloat16* shadow, size_t i, float4 v, uint32_t cols, uint32_t ld);

__device__ __forceinline__ void st_shadow_bf16x4(void* p, float4 v) {
    uint2 u;
    u.x = tfy_pack_bf16x2(v.x, v.y);
    u.y = tfy_pack_bf16x2(v.z, v.w);
    asm volatile("st.global.relaxed.sys.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(u.x), "r"(u.y) : "memory");
}

__device__ __forceinline__ void st_shadow4(__nv_bfloat16* shadow, size_t i, float4 v, uint32_t cols, uint32_t ld) {
    if (cols == 0 || ((cols & 3u) == 0 && (ld & 3u) == 0)) {
        st_shadow_bf16x4(shadow + shadow_index(i, cols, ld), v);       // the 4-pack stays inside one row
    } else {
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) shadow[shadow_index(i + k, cols, ld)] = __float2bfloat16(f[k]);
    }
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
namespace {

__device__ __forceinline__ float4 f4_mul(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void st_peer_f32x4(void* p, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float ftrl_w(float z, float n_new, float lr, float l1, float l2, float beta) {
    if (fabsf(z) <= l1) return 0.f;
    return -(z - copysignf(l1, z)) / ((beta + sqrtf(n_new)) / lr + 2.f * l2);
}

// One 4-element pack of an asynchronous update against the peer master: gv already scaled (and decayed).
// Returns the new weights when they are known (needed for the bf16 shadow), via `neww`; `has_new` tells whether
// the caller must refresh the shadow from them (true whenever a shadow exists).
__device__ __forceinline__ void ps_apply4(int opt, float* w, float* s1, float* s2, float4 gv, float lr, float eps,
                                          float p1, float p2, float p3, bool want_new, float4& neww) {
    if (opt == TFY_OPT_FTRL) {
        const float4 g2 = make_float4(gv.x * gv.x, gv.y * gv.y, gv.z * gv.z, gv.w * gv.w);
        const float4 n_old = atom_add_f32x4(s1, g2);
        const float4 n_new = make_float4(n_old.x + g2.x, n_old.y + g2.y, n_old.z + g2.z, n_old.w + g2.w);
        const float4 wv = ld_peer_f32x4(w);
        const float inv_lr = 1.f / lr;
        const float4 dz = make_float4(gv.x - (sqrtf(n_new.x) - sqrtf(n_old.x)) * inv_lr * wv.x,
                                      gv.y - (sqrtf(n_new.y) - sqrtf(n_old.y)) * inv_lr * wv.y,
                                      gv.z - (sqrtf(n_new.z) - sqrtf(n_old.z)) * inv_lr * wv.z,
                                      gv.w - (sqrtf(n_new.w) - sqrtf(n_old.w)) * inv_lr * wv.w);
        const float4 z_old = atom_add_f32x4(s2, dz);
        neww = make_float4(ftrl_w(z_old.x + dz.x, n_new.x, lr, p1, p2, p3), ftrl_w(z_old.y + dz.y, n_new.y, lr, p1, p2, p3),
                           ftrl_w(z_old.z + dz.z, n_new.z, lr, p1, p2, p3), ftrl_w(z_old.w + dz.w, n_new.w, lr, p1, p2, p3));
        st_peer_f32x4(w, neww);        // w is a pure function of (z, n): last writer wins, as in TF without locking
        return;
    }
    float4 delta;
    if (opt == TFY_OPT_ADAGRAD) {
        const float4 g2 = make_float4(gv.x * gv.x, gv.y * gv.y, gv.z * gv.z, gv.w * gv.w);
        const float4 old = atom_add_f32x4(s1, g2);
        delta.x = -lr * gv.x / (sqrtf(old.x + g2.x) + eps);
        delta.y = -lr * gv.y / (sqrtf(old.y + g2.y) + eps);
        delta.z = -lr * gv.z / (sqrtf(old.z + g2.z) + eps);
        delta.w = -lr * gv.w / (sqrtf(old.w + g2.w) + eps);
    } else if (opt == TFY_OPT_ADAM) {
        float4 m = ld_peer_f32x4(s1), v = ld_peer_f32x4(s2);
        m = make_float4(p1 * m.x + (1.f - p1) * gv.x, p1 * m.y + (1.f - p1) * gv.y, p1 * m.z + (1.f - p1) * gv.z,
                        p1 * m.w + (1.f - p1) * gv.w);
        v = make_float4(p2 * v.x + (1.f - p2) * gv.x * gv.x, p2 * v.y + (1.f - p2) * gv.y * gv.y,
                        p2 * v.z + (1.f - p2) * gv.z * gv.z, p2 * v.w + (1.f - p2) * gv.w * gv.w);
        st_peer_f32x4(s1, m);
        st_peer_f32x4(s2, v);
        delta = make_float4(-lr * m.x / (sqrtf(v.x) + eps), -lr * m.y / (sqrtf(v.y) + eps),
                            -lr * m.z / (sqrtf(v.z) + eps), -lr * m.w / (sqrtf(v.w) + eps));
    } else {
        delta = f4_mul(gv, -lr);
    }
    if (want_new) {
        const float4 old = atom_add_f32x4(w, delta);
        neww = make_float4(old.x + delta.x, old.y + delta.y, old.z + delta.z, old.w + delta.w);
    } else {
        red_add_f32x4(w, delta);
    }
}

__device__ __forceinline__ float ps_apply1(int opt, float* w, float* s1, float* s2, float g, float lr, float eps, float p1,
                                           float p2, float p3) {
    if (opt == TFY_OPT_FTRL) {
        const float n_old = atomicAdd(s1, g * g), n_new = n_old + g * g;
        const float dz = g - (sqrtf(n_new) - sqrtf(n_old)) / lr * *w;
        const float z = atomicAdd(s2, dz) + dz;
        const float nw = ftrl_w(z, n_new, lr, p1, p2, p3);
        *w = nw;
        return nw;
    }
    float delta;
    if (opt == TFY_OPT_ADAGRAD) {
        const float old = atomicAdd(s1, g * g);
        delta = -lr * g / (sqrtf(old + g * g) + eps);
    } else if (opt == TFY_OPT_ADAM) {
        const float m = p1 * *s1 + (1.f - p1) * g, v = p2 * *s2 + (1.f - p2) * g * g;
        *s1 = m; *s2 = v;
        delta = -lr * m / (sqrtf(v) + eps);
    } else {
        delta = -lr * g;
    }
    return atomicAdd(w, delta) + delta;
}

}  // namespace

// The optimizer is a per-segment (= per-variable) runtime value (uniform per blockIdx.y).  grad_scale multiplies the
// local gradient first; adam_scale[0] = sqrt(1-b2^t)/(1-b1^t) lives in device memory so that a captured CUDA graph
// sees the bias correction advance.
template <typename GT>
__global__ void __launch_bounds__(256)
tfy_ps_push_kernel(const TfyPsSeg* __restrict__ segs, float grad_scale, const float* __restrict__ adam_scale) {
    const TfyPsSeg sg = segs[blockIdx.y];
    float* w = reinterpret_cast<float*>(sg.remote_w);
    float* s1 = reinterpret_cast<float*>(sg.remote_s1);
    float* s2 = reinterpret_cast<float*>(sg.remote_s2);
    __nv_bfloat16* shadow = reinterpret_cast<__nv_bfloat16*>(sg.remote_shadow);
    const GT* g = reinterpret_cast<const GT*>(sg.local);
    const int opt = sg.opt;
    const float wd = sg.wd, eps = sg.eps, p1 = sg.p1, p2 = sg.p2, p3 = sg.p3;
    const float lr = (opt == TFY_OPT_ADAM && adam_scale) ? sg.lr * adam_scale[0] : sg.lr;
    const size_t n4 = sg.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 gv = f4_mul(ld_local4<GT>(g + i * 4), grad_scale);
        if (wd != 0.f) {
            const float4 wv = ld_peer_f32x4(w + i * 4);
            gv.x += wd * wv.x; gv.y += wd * wv.y; gv.z += wd * wv.z; gv.w += wd * wv.w;
        }
        float4 neww;
        ps_apply4(opt, w + i * 4, s1 + i * 4, s2 + i * 4, gv, lr, eps, p1, p2, p3, shadow != nullptr, neww);
        if (shadow) st_shadow4(shadow, i * 4, neww, sg.cols, sg.shadow_ld);
    }
    if (blockIdx.x == 0) {   // scalar tail (n % 4 elements)
        for (size_t i = n4 * 4 + threadIdx.x; i < sg.n; i += blockDim.x) {
            float gv = (float)g[i] * grad_scale;
            if (wd != 0.f) gv += wd * w[i];
            const float nw = ps_apply1(opt, w + i, s1 + i, s2 + i, gv, lr, eps, p1, p2, p3);
            if (shadow) shadow[shadow_index(i, sg.cols, sg.shadow_ld)] = __float2bfloat16(nw);
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

// any D (the wide tower's per-bucket weights are [V, 1] tables): one element per thread
template <typename OT>
__global__ void __launch_bounds__(256)
tfy_ps_embedding_bag_scalar_kernel(const float* __restrict__ table, const long long* __restrict__ ids,
                                   OT* __restrict__ out, int B, int L, int D, long long V, int mean) {
    const size_t total = (size_t)B * D;
    const float scale = mean ? 1.f / (float)L : 1.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / D, d = i % D;
        float acc = 0.f;
        for (int j = 0; j < L; ++j) {
            const long long id = ids[(size_t)b * L + j];
            if (id < 0 || id >= V) continue;
            float v;
            asm volatile("ld.global.relaxed.sys.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(table + (size_t)id * D + d)
                         : "memory");
            acc += v;
        }
        out[(size_t)b * D + d] = (OT)(acc * scale);
    }
}

// ------------------------------------------------------------------------------------ sparse push
// For every (b, j): row = ids[b, j]; g = dout[b, :] * (mean ? 1/L : 1) applied to the peer row.
template <typename GT>
__global__ void __launch_bounds__(256)
tfy_ps_push_rows_kernel(float* __restrict__ table, float* __restrict__ s1_table, float* __restrict__ s2_table,
                        const long long* __restrict__ ids, const GT* __restrict__ dout, int B, int L, int D, long long V,
                        int mean, TfyPsHyper h, const float* __restrict__ adam_scale, int dout_ld) {
    const int D4 = D / 4;
    const size_t total = (size_t)B * L * D4;
    const float scale = (mean ? 1.f / (float)L : 1.f) * h.grad_scale;
    const float lr = (h.opt == TFY_OPT_ADAM && adam_scale) ? h.lr * adam_scale[0] : h.lr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (i % D4) * 4;
        const size_t bj = i / D4;
        const int b = bj / L;
        const long long id = ids[bj];
        if (id < 0 || id >= V) continue;
        const float4 gv = f4_mul(ld_local4<GT>(dout + (size_t)b * dout_ld + d), scale);
        const size_t o = (size_t)id * D + d;
        float4 unused;
        ps_apply4(h.opt, table + o, s1_table + o, s2_table + o, gv, lr, h.eps, h.p1, h.p2, h.p3, false, unused);
    }
}

template <typename GT>
__global__ void __launch_bounds__(256)
tfy_ps_push_rows_scalar_kernel(float* __restrict__ table, float* __restrict__ s1_table, float* __restrict__ s2_table,
                               const long long* __restrict__ ids, const GT* __restrict__ dout, int B, int L, int D,
                               long long V, int mean, TfyPsHyper h, const float* __restrict__ adam_scale, int dout_ld) {
    const size_t total = (size_t)B * L * D;
    const float scale = (mean ? 1.f / (float)L : 1.f) * h.grad_scale;
    const float lr = (h.opt == TFY_OPT_ADAM && adam_scale) ? h.lr * adam_scale[0] : h.lr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = i % D;
        const size_t bj = i / D;
        const int b = bj / L;
        const long long id = ids[bj];
        if (id < 0 || id >= V) continue;
        const size_t o = (size_t)id * D + d;
        ps_apply1(h.opt, table + o, s1_table + o, s2_table + o, (float)dout[(size_t)b * dout_ld + d] * scale, lr, h.eps, h.p1,
                  h.p2, h.p3);
    }
}

// ------------------------------------------------------------------------------- multi-table sparse ops
// A wide-and-deep step touches one row in each of T tables per example (26 hashed categorical columns, twice: the
// wide tower's [V, 1] tables and the deep tower's [V, 64] tables).  One launch per table made the step launch-bound
// (52 gathers + 52 pushes + 26 adds of ~3 us each, profiles/r2/ps_step_profile_r2q.txt); these kernels take the
// whole group: tables = device array of T peer pointers, ids = int64 [T, B].
//   multi_bag      : out[b, d] = sum_t table_t[ids[t, b], d]                      (the wide tower's logit)
//   multi_push_rows: row ids[t, b] of table t receives dout[b, t * col_stride + d] (col_stride 0: every table gets
//                    the same gradient -- wide tower; col_stride D: column slices of a wider matrix -- the fused
//                    first deep layer's dx), optimizer fused as in tfy_ps_push_rows.
template <typename OT>
__global__ void __launch_bounds__(256)
tfy_ps_multi_bag_kernel(const uint64_t* __restrict__ tables, const long long* __restrict__ ids, OT* __restrict__ out,
                        int B, int T, int D, long long V) {
    const size_t total = (size_t)B * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / D, d = i % D;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const long long id = ids[(size_t)t * B + b];
            if (id < 0 || id >= V) continue;
            float v;
            asm volatile("ld.global.relaxed.sys.L1::no_allocate.f32 %0, [%1];"
                         : "=f"(v)
                         : "l"(reinterpret_cast<const float*>(tables[t]) + (size_t)id * D + d)
                         : "memory");
            acc += v;
        }
        out[i] = (OT)acc;
    }
}

template <typename GT, int VEC>
__global__ void __launch_bounds__(256)
tfy_ps_multi_push_rows_kernel(const uint64_t* __restrict__ tables, const uint64_t* __restrict__ s1s,
                              const uint64_t* __restrict__ s2s, const long long* __restrict__ ids,
                              const GT* __restrict__ dout, int B, int T, int D, long long V, TfyPsHyper h,
                              const float* __restrict__ adam_scale, int dout_ld, int col_stride) {
    const int DV = D / VEC;
    const size_t total = (size_t)T * B * DV;
    const float lr = (h.opt == TFY_OPT_ADAM && adam_scale) ? h.lr * adam_scale[0] : h.lr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (i % DV) * VEC;
        const size_t tb = i / DV;
        const int b = tb % B, t = tb / B;
        const long long id = ids[tb];
        if (id < 0 || id >= V) continue;
        const size_t o = (size_t)id * D + d;
        float* w = reinterpret_cast<float*>(tables[t]) + o;
        float* s1 = s1s ? reinterpret_cast<float*>(s1s[t]) + o : nullptr;
        float* s2 = s2s ? reinterpret_cast<float*>(s2s[t]) + o : nullptr;
        const GT* g = dout + (size_t)b * dout_ld + (size_t)t * col_stride + d;
        if (VEC == 4) {
            float4 unused;
            ps_apply4(h.opt, w, s1, s2, f4_mul(ld_local4<GT>(g), h.grad_scale), lr, h.eps, h.p1, h.p2, h.p3, false, unused);
        } else {
            ps_apply1(h.opt, w, s1, s2, (float)g[0] * h.grad_scale, lr, h.eps, h.p1, h.p2, h.p3);
        }
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
        st_shadow4(sh, i * 4, ld_peer_f32x4(w + i * 4), sg.cols, sg.shadow_ld);
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < sg.n; i += blockDim.x)
            sh[shadow_index(i, sg.cols, sg.shadow_ld)] = __float2bfloat16(w[i]);
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

// adam_scale: device pointer to sqrt(1-b2^t)/(1-b1^t) (may be nullptr when no segment uses Adam)
int tfy_ps_push(const TfyPsSeg* segs, int nseg, size_t max_n, int grad_is_bf16, float grad_scale,
                const float* adam_scale, cudaStream_t s) {
    if (nseg <= 0) return 0;
    size_t gx = (max_n / 4 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid((unsigned)gx, nseg);
    if (grad_is_bf16) tfy_ps_push_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(segs, grad_scale, adam_scale);
    else tfy_ps_push_kernel<float><<<grid, 256, 0, s>>>(segs, grad_scale, adam_scale);
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
    if (D % 4) {
        size_t g1 = ((size_t)B * D + 255) / 256;
        if (g1 < 1) g1 = 1;
        if (g1 > 148 * 8) g1 = 148 * 8;
        if (out_is_bf16)
            tfy_ps_embedding_bag_scalar_kernel<__nv_bfloat16><<<(unsigned)g1, 256, 0, s>>>(
                (const float*)table, (const long long*)ids, (__nv_bfloat16*)out, B, L, D, V, mean);
        else
            tfy_ps_embedding_bag_scalar_kernel<float><<<(unsigned)g1, 256, 0, s>>>(
                (const float*)table, (const long long*)ids, (float*)out, B, L, D, V, mean);
        return (int)cudaGetLastError();
    }
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

int tfy_ps_push_rows(void* table, void* s1_table, void* s2_table, const void* ids, const void* dout, int grad_is_bf16,
                     int B, int L, int D, long long V, int mean, int opt, float lr, float eps, float p1, float p2, float p3,
                     float grad_scale, const float* adam_scale, int dout_ld, cudaStream_t s) {
    // dout_ld: row pitch (elements) of dout; 0 = D (contiguous [B, D]).  A larger pitch lets the caller push a
    // column slice of a wider gradient matrix (the fused first layer's dx) without copying it out.
    if (dout_ld <= 0) dout_ld = D;
    if (opt != TFY_OPT_SGD && opt != TFY_OPT_ADAGRAD && opt != TFY_OPT_ADAM && opt != TFY_OPT_FTRL) return -3;
    if (opt != TFY_OPT_SGD && !s1_table) return -4;
    if ((opt == TFY_OPT_ADAM || opt == TFY_OPT_FTRL) && !s2_table) return -4;
    size_t gx = ((size_t)B * L * (D / 4) + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
    TfyPsHyper h;
    h.opt = opt; h.lr = lr; h.eps = eps; h.wd = 0.f; h.p1 = p1; h.p2 = p2; h.p3 = p3; h.grad_scale = grad_scale;
    if (D % 4) {
        size_t g1 = ((size_t)B * L * D + 255) / 256;
        if (g1 < 1) g1 = 1;
        if (g1 > 148 * 8) g1 = 148 * 8;
        if (grad_is_bf16)
            tfy_ps_push_rows_scalar_kernel<__nv_bfloat16><<<(unsigned)g1, 256, 0, s>>>(
                (float*)table, (float*)s1_table, (float*)s2_table, (const long long*)ids, (const __nv_bfloat16*)dout, B,
                L, D, V, mean, h, adam_scale, dout_ld);
        else
            tfy_ps_push_rows_scalar_kernel<float><<<(unsigned)g1, 256, 0, s>>>(
                (float*)table, (float*)s1_table, (float*)s2_table, (const long long*)ids, (const float*)dout, B, L, D, V,
                mean, h, adam_scale, dout_ld);
        return (int)cudaGetLastError();
    }
    if (grad_is_bf16)
        tfy_ps_push_rows_kernel<__nv_bfloat16><<<(unsigned)gx, 256, 0, s>>>(
            (float*)table, (float*)s1_table, (float*)s2_table, (const long long*)ids, (const __nv_bfloat16*)dout, B, L, D,
            V, mean, h, adam_scale, dout_ld);
    else
        tfy_ps_push_rows_kernel<float><<<(unsigned)gx, 256, 0, s>>>((float*)table, (float*)s1_table, (float*)s2_table,
                                                                    (const long long*)ids, (const float*)dout, B, L, D,
                                                                    V, mean, h, adam_scale, dout_ld);
    return (int)cudaGetLastError();
}

int tfy_ps_multi_bag(const uint64_t* tables, const void* ids, void* out, int out_is_bf16, int B, int T, int D,
                     long long V, cudaStream_t s) {
    if (B < 1 || T < 1 || D < 1) return -2;
    size_t gx = ((size_t)B * D + 255) / 256;
    if (gx > 148 * 8) gx = 148 * 8;
    if (out_is_bf16)
        tfy_ps_multi_bag_kernel<__nv_bfloat16><<<(unsigned)gx, 256, 0, s>>>(tables, (const long long*)ids,
                                                                           (__nv_bfloat16*)out, B, T, D, V);
    else
        tfy_ps_multi_bag_kernel<float><<<(unsigned)gx, 256, 0, s>>>(tables, (const long long*)ids, (float*)out, B, T, D, V);
    return (int)cudaGetLastError();
}

// s1s / s2s: device arrays of T slot-table pointers (nullptr when the optimizer has no such slot)
int tfy_ps_multi_push_rows(const uint64_t* tables, const uint64_t* s1s, const uint64_t* s2s, const void* ids,
                           const void* dout, int grad_is_bf16, int B, int T, int D, long long V, int opt, float lr,
                           float eps, float p1, float p2, float p3, float grad_scale, const float* adam_scale,
                           int dout_ld, int col_stride, cudaStream_t s) {
    if (opt != TFY_OPT_SGD && opt != TFY_OPT_ADAGRAD && opt != TFY_OPT_ADAM && opt != TFY_OPT_FTRL) return -3;
    if (opt != TFY_OPT_SGD && !s1s) return -4;
    if ((opt == TFY_OPT_ADAM || opt == TFY_OPT_FTRL) && !s2s) return -4;
    if (dout_ld <= 0) dout_ld = D;
    TfyPsHyper h;
    h.opt = opt; h.lr = lr; h.eps = eps; h.wd = 0.f; h.p1 = p1; h.p2 = p2; h.p3 = p3; h.grad_scale = grad_scale;
    const bool vec = (D % 4 == 0) && (dout_ld % 4 == 0) && (col_stride % 4 == 0);
    size_t gx = ((size_t)T * B * (vec ? D / 4 : D) + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 148 * 8) gx = 148 * 8;
#define TFY_MPR(GT, VV)                                                                                              \
    tfy_ps_multi_push_rows_kernel<GT, VV><<<(unsigned)gx, 256, 0, s>>>(tables, s1s, s2s, (const long long*)ids,       \
                                                                       (const GT*)dout, B, T, D, V, h, adam_scale,    \
                                                                       dout_ld, col_stride)
    if (grad_is_bf16) { if (vec) TFY_MPR(__nv_bfloat16, 4); else TFY_MPR(__nv_bfloat16, 1); }
    else { if (vec) TFY_MPR(float, 4); else TFY_MPR(float, 1); }
#undef TFY_MPR
    return (int)cudaGetLastError();
}

}  // extern "C"
