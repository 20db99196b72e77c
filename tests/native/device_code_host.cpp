// Host build of the fused step's per-element update (ops/csrc/tfy_fused_step.cuh: tfy_step_consts + tfy_opt_update_rt),
// so that the formulas the K4 kernel applies -- SGD-momentum / Adadelta / Adam(W) / Adagrad / FTRL, bias correction
// from the device-resident step counter included -- can be checked against torch.optim on a CPU-only box.
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }

#include "tfy_fused_step.cuh"

extern "C" void tfy_host_opt_step(int opt, TfyOptHyper* hp, int world, float* p, const float* g, float* s1, float* s2,
                                  int n) {
    const TfyStepConsts k = tfy_step_consts(hp, world, opt);
    for (int i = 0; i < n; ++i) {
        float pi = p[i], a = s1[i], b = s2[i];
        tfy_opt_update_rt<-1>(opt, pi, g[i] * k.gscale, a, b, k.lr, k.p1, k.p2, k.eps, k.wd, k.flags, k.lr_bc1,
                              k.bc2_rsqrt, k.first_step);
        p[i] = pi; s1[i] = a; s2[i] = b;
    }
    hp->step += 1;
}

// the counter-based dropout generator shared by the forward kernels and the mask-recomputing backward kernels
extern "C" void tfy_host_uniform(uint32_t seed, uint32_t step, uint64_t idx0, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = tfy_uniform(seed, step, idx0 + (uint64_t)i);
}

// fp32 -> bf16 pack (round to nearest even) and back, as the fused step stores parameters and reads gradients
extern "C" void tfy_host_bf16_roundtrip(const float* in, float* out, int n8) {
    for (int i = 0; i < n8; ++i) {
        const uint4 v = TfyPack<__nv_bfloat16>::pack(in + 8 * i);
        TfyPack<__nv_bfloat16>::unpack(v, out + 8 * i);
    }
}

// patch scheduling of the persistent convolution kernels: the (image pair, tile row, tile column) sequence of one CTA
#include "tfy_conv_index.cuh"
extern "C" void tfy_host_patch_walk(int tiles_x, int tiles_y, int n_cta, int block, int grid, int explicit_n_cta,
                                    int count, int* out) {
    blockIdx.x = (unsigned)block;
    gridDim.x = (unsigned)grid;
    CPatchIter it = explicit_n_cta ? CPatchIter(tiles_x, tiles_y, n_cta) : CPatchIter(tiles_x, tiles_y);
    for (int i = 0; i < count; ++i) {
        out[3 * i] = it.bz; out[3 * i + 1] = it.ty; out[3 * i + 2] = it.tx;
        it.next();
    }
}
