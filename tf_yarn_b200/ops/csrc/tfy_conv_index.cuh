// Patch scheduling of the persistent convolution kernels (included by tfy_conv.cu; kept in its own header so that the
// index arithmetic can be unit-tested on the host: tests/test_native_device_code.py builds it with g++).
#pragma once

// Walks the patch list of a persistent CTA: p = blockIdx.x + i * gridDim.x decomposed into (image pair,
// tile row, tile column) incrementally -- no integer division per patch in the hot loops.
struct CPatchIter {
    int tx, ty, bz, sx, sy, sb, tiles_x, tiles_y;
    // n_cta: CTAs sharing the patch list (the whole grid unless the kernel carries extra communication CTAs)
    __device__ __forceinline__ CPatchIter(int tiles_x_, int tiles_y_, int n_cta = (int)gridDim.x)
        : tiles_x(tiles_x_), tiles_y(tiles_y_) {
        const int tiles = tiles_x * tiles_y;
        int p = (int)blockIdx.x;
        bz = p / tiles; p -= bz * tiles; ty = p / tiles_x; tx = p - ty * tiles_x;
        int q = n_cta;
        sb = q / tiles; q -= sb * tiles; sy = q / tiles_x; sx = q - sy * tiles_x;
    }
    __device__ __forceinline__ void next() {
        tx += sx;
        if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += sy;
        if (ty >= tiles_y) { ty -= tiles_y; ++bz; }
        bz += sb;
    }
};
