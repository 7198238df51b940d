// densify.cu -- densification statistics of one training view in ONE pass over the Gaussians.
//
// Replaces the PyTorch statements that consume the rasterizer's outputs right after the backward pass
// (wildgaussians/method.py:1997-1998 and GaussianModel.add_densification_stats, :1470-1477):
//     max_radii2D[vis]                = max(max_radii2D[vis], radii[vis])
//     xyz_grad[vis]                  += |viewspace_grad[vis, :2]|
//     xyz_gradient_accum_abs[vis]    += |viewspace_grad[vis, 2:]|            (use_gof_abs_gradient)
//     xyz_gradient_accum_abs_max[vis] = max(., |viewspace_grad[vis, 2:]|)
//     denom[vis]                     += 1
// with vis = radii > 0: five boolean-indexed gathers / scatters (each a nonzero() + index kernels + a host sync for
// the index count) become one bandwidth-bound kernel reading 16 B and touching 20 B per Gaussian.
#include "common.cuh"

namespace gsr {

__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const int* __restrict__ radii,
                                                            const float* __restrict__ viewspace_grad,   // [P,3]
                                                            float* __restrict__ max_radii2D, float* __restrict__ xyz_grad,
                                                            float* __restrict__ accum_abs, float* __restrict__ accum_abs_max,
                                                            float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = viewspace_grad[3 * (size_t)i], gy = viewspace_grad[3 * (size_t)i + 1], gz = viewspace_grad[3 * (size_t)i + 2];
    max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    xyz_grad[i] += sqrtf(gx * gx + gy * gy);               // torch.norm(grad[:, :2], dim=-1)
    if (accum_abs) {
        const float a = fabsf(gz);                          // torch.norm over the single column grad[:, 2:]
        accum_abs[i] += a;
        accum_abs_max[i] = fmaxf(accum_abs_max[i], a);
    }
    denom[i] += 1.0f;
}

}  // namespace gsr

extern "C" int gsr_densification_stats(int P, const int* radii, const float* viewspace_grad, float* max_radii2D, float* xyz_grad,
                                       float* xyz_gradient_accum_abs, float* xyz_gradient_accum_abs_max, float* denom, void* stream) {
    using namespace gsr;
    if (P < 0) { set_error("bad P"); return GSR_E_INVALID; }
    if (P == 0) return 0;
    if (!radii || !viewspace_grad || !max_radii2D || !xyz_grad || !denom || ((xyz_gradient_accum_abs == nullptr) != (xyz_gradient_accum_abs_max == nullptr))) {
        set_error("a required pointer is NULL");
        return GSR_E_INVALID;
    }
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, radii, viewspace_grad, max_radii2D, xyz_grad,
                                                                            xyz_gradient_accum_abs, xyz_gradient_accum_abs_max, denom);
    count_launches(1);
    GSR_CUDA(cudaGetLastError());
    return 0;
}
