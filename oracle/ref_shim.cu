// ref_shim.cu -- TEST INFRASTRUCTURE, not product code.
//
// A C-ABI wrapper around the UNMODIFIED reference rasterizer core
// (submodules/diff-gaussian-rasterization/cuda_rasterizer/{rasterizer_impl,forward,backward}.cu),
// compiled from the sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libdgr_ref.so.  No reference source is copied into this repository: this file only
// includes the reference's public header (cuda_rasterizer/rasterizer.h) through the -I path and
// calls CudaRasterizer::Rasterizer::{forward,backward,markVisible}.
//
// It exists so that the GPU parity tests and `bench.py --impl reference` can run the real
// reference on the GPU box (where /root/reference does not exist but the built .so travels).
// Only tests/, __graft_entry__.smoke() and bench.py may load it.
#include <cstddef>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "cuda_rasterizer/rasterizer.h"

extern "C" {

typedef char* (*ref_alloc_fn)(void* ctx, int which, size_t bytes);  // which: 0 geom, 1 binning, 2 img

static thread_local std::string g_ref_err;
const char* ref_last_error(void) { return g_ref_err.c_str(); }

// CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-340); returns num_rendered or -1
int ref_forward(ref_alloc_fn alloc, void* ctx, int P, int D, int M, const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                float tan_fovy, float kernel_size, const float* subpixel_offset, int prefiltered, float* out_color,
                int* radii, int debug) {
    try {
        std::function<char*(size_t)> geomF = [=](size_t n) { return alloc(ctx, 0, n); };
        std::function<char*(size_t)> binF = [=](size_t n) { return alloc(ctx, 1, n); };
        std::function<char*(size_t)> imgF = [=](size_t n) { return alloc(ctx, 2, n); };
        return CudaRasterizer::Rasterizer::forward(geomF, binF, imgF, P, D, M, background, W, H, means3D, shs,
                                                   colors_precomp, opacities, scales, scale_modifier, rotations,
                                                   cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
                                                   tan_fovy, kernel_size, subpixel_offset, prefiltered != 0,
                                                   out_color, radii, debug != 0);
    } catch (const std::exception& e) {
        g_ref_err = e.what();
        return -1;
    }
}

// CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:344-444); all dL_* must be zero-filled
int ref_backward(int P, int D, int M, int R, const float* background, int W, int H, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, float kernel_size,
                 const float* subpixel_offset, const int* radii, char* geom_buffer, char* binning_buffer,
                 char* img_buffer, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                 float* dL_drot, int debug) {
    try {
        CudaRasterizer::Rasterizer::backward(P, D, M, R, background, W, H, means3D, shs, colors_precomp, scales,
                                             scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                                             campos, tan_fovx, tan_fovy, kernel_size, subpixel_offset, radii,
                                             geom_buffer, binning_buffer, img_buffer, dL_dpix, dL_dmean2D, dL_dconic,
                                             dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale,
                                             dL_drot, debug != 0);
        return 0;
    } catch (const std::exception& e) {
        g_ref_err = e.what();
        return -1;
    }
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return 0;
}

}  // extern "C"
