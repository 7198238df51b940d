// sh_math.cuh -- real spherical-harmonics basis (degrees 0..3) and its gradient, shared by the per-Gaussian backward of
// the rasterizer (preprocess_bwd.cu: backward of the in-kernel SH colours, reference backward.cu:20-139) and by the fused
// colour op (appearance.cu: method.py:493-548).  The colour of a channel is  sum_k b_k(dir) * coeff[k];  its gradient
// w.r.t. the coefficients is the basis value, w.r.t. the direction  sum_k coeff_weight[k] * grad b_k(dir).
// (The FORWARD evaluation inside the rasterizer lives in preprocess_fwd.cu with the reference's exact rounding sequence;
// this header is for gradients and for the colour op, where only fp32 accuracy matters.)
#pragma once
#include <cuda_runtime.h>

namespace gsr {

constexpr float SHK0 = 0.28209479177387814f, SHK1 = 0.4886025119029199f;
constexpr float SHK2_0 = 1.0925484305920792f, SHK2_1 = -1.0925484305920792f, SHK2_2 = 0.31539156525252005f,
                SHK2_3 = -1.0925484305920792f, SHK2_4 = 0.5462742152960396f;
constexpr float SHK3_0 = -0.5900435899266435f, SHK3_1 = 2.890611442640554f, SHK3_2 = -0.4570457994644658f,
                SHK3_3 = 0.3731763325901154f, SHK3_4 = -0.4570457994644658f, SHK3_5 = 1.445305721320277f,
                SHK3_6 = -0.5900435899266435f;

// basis values for the active degree; entries above the degree are 0
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float (&b)[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) b[k] = 0.f;
    b[0] = SHK0;
    if (deg > 0) {
        b[1] = -SHK1 * y; b[2] = SHK1 * z; b[3] = -SHK1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = SHK2_0 * xy; b[5] = SHK2_1 * yz; b[6] = SHK2_2 * (2.f * zz - xx - yy);
            b[7] = SHK2_3 * xz; b[8] = SHK2_4 * (xx - yy);
            if (deg > 2) {
                b[9] = SHK3_0 * y * (3.f * xx - yy); b[10] = SHK3_1 * xy * z;
                b[11] = SHK3_2 * y * (4.f * zz - xx - yy); b[12] = SHK3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
                b[13] = SHK3_4 * x * (4.f * zz - xx - yy); b[14] = SHK3_5 * z * (xx - yy);
                b[15] = SHK3_6 * x * (xx - 3.f * yy);
            }
        }
    }
}

// sum_k g[k] * grad b_k(x, y, z)   (x, y, z treated as independent variables, like autograd on the polynomial)
__device__ __forceinline__ float3 sh_basis_grad_dot(int deg, float x, float y, float z, const float (&g)[16]) {
    float3 r = {0.f, 0.f, 0.f};
    if (deg > 0) {
        r.y += -SHK1 * g[1]; r.z += SHK1 * g[2]; r.x += -SHK1 * g[3];
        if (deg > 1) {
            r.x += SHK2_0 * y * g[4] - 2.f * SHK2_2 * x * g[6] + SHK2_3 * z * g[7] + 2.f * SHK2_4 * x * g[8];
            r.y += SHK2_0 * x * g[4] + SHK2_1 * z * g[5] - 2.f * SHK2_2 * y * g[6] - 2.f * SHK2_4 * y * g[8];
            r.z += SHK2_1 * y * g[5] + 4.f * SHK2_2 * z * g[6] + SHK2_3 * x * g[7];
            if (deg > 2) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r.x += SHK3_0 * 6.f * xy * g[9] + SHK3_1 * yz * g[10] - SHK3_2 * 2.f * xy * g[11] - SHK3_3 * 6.f * xz * g[12] +
                       SHK3_4 * (4.f * zz - 3.f * xx - yy) * g[13] + SHK3_5 * 2.f * xz * g[14] + SHK3_6 * 3.f * (xx - yy) * g[15];
                r.y += SHK3_0 * 3.f * (xx - yy) * g[9] + SHK3_1 * xz * g[10] + SHK3_2 * (4.f * zz - xx - 3.f * yy) * g[11] -
                       SHK3_3 * 6.f * yz * g[12] - SHK3_4 * 2.f * xy * g[13] - SHK3_5 * 2.f * yz * g[14] - SHK3_6 * 6.f * xy * g[15];
                r.z += SHK3_1 * xy * g[10] + SHK3_2 * 8.f * yz * g[11] + SHK3_3 * (6.f * zz - 3.f * xx - 3.f * yy) * g[12] +
                       SHK3_4 * 8.f * xz * g[13] + SHK3_5 * (xx - yy) * g[14];
            }
        }
    }
    return r;
}

// d(v / |v|) applied to a gradient w.r.t. the unit vector: (g - dir (dir . g)) / |v|
__device__ __forceinline__ float3 through_normalize(float3 dir, float inv_len, float3 g) {
    const float d = g.x * dir.x + g.y * dir.y + g.z * dir.z;
    return {(g.x - dir.x * d) * inv_len, (g.y - dir.y * d) * inv_len, (g.z - dir.z * d) * inv_len};
}

}  // namespace gsr
