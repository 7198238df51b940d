// preprocess_bwd.cu -- per-Gaussian backward: from the compositing partials
// (dL/d{mean2D, conic, opacity*coef, colour}) to the gradients of the rasterizer inputs.
//
// One fused kernel replaces the reference's two per-Gaussian launches, computeCov2DCUDA
// (backward.cu:144-310) and preprocessCUDA<3> (backward.cu:382-432, with the SH backward
// :20-139 and the scale/rotation backward :314-377).  Fusing removes the dL_dcov3D and
// dL_dmeans round trips through HBM between the two, and cov3D is recomputed from
// scale/rotation instead of being read back from the forward's scratch.  Every output row
// is written here (zeros for Gaussians that were not rendered), so the host does not have
// to zero-fill nine gradient tensors first (rasterize_points.cu:157-165).
// The reference's opacity-compensation gradient (backward.cu:199-217) is mixed fp32/fp64; it is evaluated in
// fp32 here (gradients are compared at 1e-3; see the comment in the kernel).
#include "common.cuh"
#include "gaussian_math.cuh"
#include <cstdlib>

namespace gsr {

struct PreBwdParams {
    int P, D, M;
    int sh_vec;             // SH rows are 16-byte aligned multiples of 16 bytes
    float focal_x, focal_y, tan_fovx, tan_fovy, kernel_size, scale_modifier;
    const float* means3D;
    const int* radii;
    const float* shs;
    const uint8_t* clamped;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* view;
    const float* proj;
    const float* campos;
    const float4* rec;
    const float* accum;     // [P][12]
    float* dL_dmean2D;      // [P,3]
    float* dL_dconic;       // [P,4] or NULL
    float* dL_dopacity;     // [P]
    float* dL_dcolor;       // [P,3]
    float* dL_dmean3D;      // [P,3]
    float* dL_dcov3D;       // [P,6] or NULL
    float* dL_dsh;          // [P,M,3] or NULL
    float* dL_dscale;       // [P,3] or NULL
    float* dL_drot;         // [P,4] or NULL
};

// direction-normalisation Jacobian applied to dv (auxiliary.h:107-117)
__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// SH colour backward (backward.cu:20-139): writes the dL_dsh row, returns the mean gradient caused by the view
// direction.  `vec`: rows are aligned multiples of 16 bytes -- the coefficients are read and the gradient row is
// written with 128-bit accesses (a scalar access pattern makes every instruction of a warp touch 32 sectors, one per
// Gaussian, and the kernel becomes bound by L1/L2 sector traffic: measured 8x slower per Gaussian).
__device__ __forceinline__ float3 sh_backward(int deg, int max_coeffs, const float* __restrict__ sh_base, bool vec, float3 mean,
                                              const float* __restrict__ campos, unsigned clamp_bits, V3 dL_dRGB,
                                              float* __restrict__ dL_dsh_base) {
    V3 pos = {mean.x, mean.y, mean.z};
    V3 cam = {campos[0], campos[1], campos[2]};
    V3 dir_orig = pos - cam;
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};

    dL_dRGB.x *= (clamp_bits & 1u) ? 0 : 1;
    dL_dRGB.y *= (clamp_bits & 2u) ? 0 : 1;
    dL_dRGB.z *= (clamp_bits & 4u) ? 0 : 1;

    const int used = (deg + 1) * (deg + 1);
    float shl[48];
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(sh_base);
        const int n4 = (3 * used + 3) >> 2;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            if (i < n4) {
                const float4 v = __ldg(p4 + i);
                shl[4 * i + 0] = v.x; shl[4 * i + 1] = v.y; shl[4 * i + 2] = v.z; shl[4 * i + 3] = v.w;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 48; ++i)
            if (i < 3 * used) shl[i] = sh_base[i];
    }

    V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
    const float x = dir.x, y = dir.y, z = dir.z;

    auto sh = [&](int k) { return V3{shl[3 * k], shl[3 * k + 1], shl[3 * k + 2]}; };
    // basis weight of every coefficient (0 above the active degree: those receive no gradient)
    float w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0.f;

    w[0] = GSR_SH_C0;
    if (deg > 0) {
        w[1] = -GSR_SH_C1 * y;
        w[2] = GSR_SH_C1 * z;
        w[3] = -GSR_SH_C1 * x;

        dRGBdx = -GSR_SH_C1 * sh(3);
        dRGBdy = -GSR_SH_C1 * sh(1);
        dRGBdz = GSR_SH_C1 * sh(2);

        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
            w[4] = GSR_SH_C2_0 * xy;
            w[5] = GSR_SH_C2_1 * yz;
            w[6] = GSR_SH_C2_2 * (2.f * zz - xx - yy);
            w[7] = GSR_SH_C2_3 * xz;
            w[8] = GSR_SH_C2_4 * (xx - yy);

            dRGBdx += GSR_SH_C2_0 * y * sh(4) + GSR_SH_C2_2 * 2.f * -x * sh(6) + GSR_SH_C2_3 * z * sh(7) + GSR_SH_C2_4 * 2.f * x * sh(8);
            dRGBdy += GSR_SH_C2_0 * x * sh(4) + GSR_SH_C2_1 * z * sh(5) + GSR_SH_C2_2 * 2.f * -y * sh(6) + GSR_SH_C2_4 * 2.f * -y * sh(8);
            dRGBdz += GSR_SH_C2_1 * y * sh(5) + GSR_SH_C2_2 * 2.f * 2.f * z * sh(6) + GSR_SH_C2_3 * x * sh(7);

            if (deg > 2) {
                w[9] = GSR_SH_C3_0 * y * (3.f * xx - yy);
                w[10] = GSR_SH_C3_1 * xy * z;
                w[11] = GSR_SH_C3_2 * y * (4.f * zz - xx - yy);
                w[12] = GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
                w[13] = GSR_SH_C3_4 * x * (4.f * zz - xx - yy);
                w[14] = GSR_SH_C3_5 * z * (xx - yy);
                w[15] = GSR_SH_C3_6 * x * (xx - 3.f * yy);

                dRGBdx += (GSR_SH_C3_0 * sh(9) * 3.f * 2.f * xy +
                           GSR_SH_C3_1 * sh(10) * yz +
                           GSR_SH_C3_2 * sh(11) * -2.f * xy +
                           GSR_SH_C3_3 * sh(12) * -3.f * 2.f * xz +
                           GSR_SH_C3_4 * sh(13) * (-3.f * xx + 4.f * zz - yy) +
                           GSR_SH_C3_5 * sh(14) * 2.f * xz +
                           GSR_SH_C3_6 * sh(15) * 3.f * (xx - yy));
                dRGBdy += (GSR_SH_C3_0 * sh(9) * 3.f * (xx - yy) +
                           GSR_SH_C3_1 * sh(10) * xz +
                           GSR_SH_C3_2 * sh(11) * (-3.f * yy + 4.f * zz - xx) +
                           GSR_SH_C3_3 * sh(12) * -3.f * 2.f * yz +
                           GSR_SH_C3_4 * sh(13) * -2.f * xy +
                           GSR_SH_C3_5 * sh(14) * -2.f * yz +
                           GSR_SH_C3_6 * sh(15) * -3.f * 2.f * xy);
                dRGBdz += (GSR_SH_C3_1 * sh(10) * xy +
                           GSR_SH_C3_2 * sh(11) * 4.f * 2.f * yz +
                           GSR_SH_C3_3 * sh(12) * 3.f * (2.f * zz - xx - yy) +
                           GSR_SH_C3_4 * sh(13) * 4.f * 2.f * xz +
                           GSR_SH_C3_5 * sh(14) * (xx - yy));
            }
        }
    }
    // dL_dsh[k][c] = w[k] * dL_dRGB[c]; coefficients above the active degree get 0 (w[k] = 0)
    if (vec) {
        float4* o4 = reinterpret_cast<float4*>(dL_dsh_base);
        const int groups = (max_coeffs * 3) >> 2;   // 4 coefficients (12 floats) = 3 float4 per group of four
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
            if (3 * g4 < groups) {
                const float w0 = w[4 * g4], w1 = w[4 * g4 + 1], w2 = w[4 * g4 + 2], w3 = w[4 * g4 + 3];
                o4[3 * g4 + 0] = make_float4(w0 * dL_dRGB.x, w0 * dL_dRGB.y, w0 * dL_dRGB.z, w1 * dL_dRGB.x);
                o4[3 * g4 + 1] = make_float4(w1 * dL_dRGB.y, w1 * dL_dRGB.z, w2 * dL_dRGB.x, w2 * dL_dRGB.y);
                o4[3 * g4 + 2] = make_float4(w2 * dL_dRGB.z, w3 * dL_dRGB.x, w3 * dL_dRGB.y, w3 * dL_dRGB.z);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < max_coeffs) {
                dL_dsh_base[3 * k + 0] = w[k] * dL_dRGB.x;
                dL_dsh_base[3 * k + 1] = w[k] * dL_dRGB.y;
                dL_dsh_base[3 * k + 2] = w[k] * dL_dRGB.z;
            }
        }
        for (int k = 16; k < max_coeffs; ++k) {      // tensors holding more than degree-3 coefficients
            dL_dsh_base[3 * k + 0] = 0.f;
            dL_dsh_base[3 * k + 1] = 0.f;
            dL_dsh_base[3 * k + 2] = 0.f;
        }
    }

    const float3 dL_ddir = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
    return dnormvdv3(float3{dir_orig.x, dir_orig.y, dir_orig.z}, dL_ddir);
}

template <int MIN_CTAS, bool HAS_SH>
__global__ void __launch_bounds__(256, MIN_CTAS) preprocess_bwd_kernel(const __grid_constant__ PreBwdParams p) {
    __shared__ float s_view[16];
    __shared__ float s_proj[16];
    if (threadIdx.x < 16) s_view[threadIdx.x] = p.view[threadIdx.x];
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = p.proj[threadIdx.x - 16];
    __syncthreads();

    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    const bool rendered = p.radii[idx] > 0;

    float o_mean2D[3] = {0.f, 0.f, 0.f};
    float o_conic[4] = {0.f, 0.f, 0.f, 0.f};
    float o_opacity = 0.f;
    float o_color[3] = {0.f, 0.f, 0.f};
    float o_mean3D[3] = {0.f, 0.f, 0.f};
    float o_cov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float o_scale[3] = {0.f, 0.f, 0.f};
    float o_rot[4] = {0.f, 0.f, 0.f, 0.f};
    bool sh_written = false;

    if (rendered) {
        const float4* acc = reinterpret_cast<const float4*>(p.accum + (size_t)idx * 12);
        const float4 a0 = acc[0], a1 = acc[1], a2 = acc[2];
        o_mean2D[0] = a0.x; o_mean2D[1] = a0.y; o_mean2D[2] = a0.z;
        o_conic[0] = a0.w; o_conic[1] = a1.x; o_conic[3] = a1.y;
        float dL_dopacity = a1.z;
        o_color[0] = a1.w; o_color[1] = a2.x; o_color[2] = a2.y;

        const float3 mean = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};

        float cov3D_local[6];
        const float* cov3D;
        float3 sc = {0.f, 0.f, 0.f};
        float4 rot = {0.f, 0.f, 0.f, 0.f};
        if (p.cov3D_precomp != nullptr) {
            cov3D = p.cov3D_precomp + 6 * (size_t)idx;
        } else {
            sc = float3{p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]};
            rot = reinterpret_cast<const float4*>(p.rotations)[idx];
            cov3d_from_scale_rot(sc, p.scale_modifier, rot, cov3D_local);
            cov3D = cov3D_local;
        }

        // ---- 2D covariance / conic backward (backward.cu:144-310) ----
        const float3 dL_dconic = {o_conic[0], o_conic[1], o_conic[3]};
        const float4 rb = p.rec[2 * (size_t)idx + 1];
        const float combined_opacity = rb.w;
        const float h_x = p.focal_x, h_y = p.focal_y;

        const Ewa e = ewa_project(mean, h_x, h_y, p.tan_fovx, p.tan_fovy, cov3D, s_view);
        const float3 t = e.t;
        const float x_grad_mul = e.txtz < -e.limx || e.txtz > e.limx ? 0 : 1;
        const float y_grad_mul = e.tytz < -e.limy || e.tytz > e.limy ? 0 : 1;
        const Mat3& T = e.T;
        const Mat3& Vrk = e.Vrk;
        Mat3 cov2D = e.cov;
        const float kernel_size = p.kernel_size;

        // Opacity-compensation gradient (backward.cu:199-217).  The reference evaluates this block in double
        // because of its double literals; only gradients (compared at 1e-3) depend on it, so it is evaluated in
        // fp32 here -- about 350 instructions of fp64 division / square root per Gaussian less.  The two clamps
        // and the det <= 1e-6 tests below select exactly the same branch as the double comparisons: no float lies
        // strictly between (float)1e-6 and 1e-6.
        const float det_0 = fmaxf(1e-6f, cov2D.m[0][0] * cov2D.m[1][1] - cov2D.m[0][1] * cov2D.m[0][1]);
        const float det_1 = fmaxf(1e-6f, (cov2D.m[0][0] + kernel_size) * (cov2D.m[1][1] + kernel_size) - cov2D.m[0][1] * cov2D.m[0][1]);
        const float inv_det1e = 1.0f / (det_1 + 1e-6f);
        const float coef = sqrtf(det_0 * inv_det1e + 1e-6f);

        const float inv_coefe = 1.0f / (coef + 1e-6f);
        const float opacity = combined_opacity * inv_coefe;
        const float dL_dcoef = dL_dopacity * opacity;
        const float dL_dsqrtcoef = dL_dcoef * 0.5f * inv_coefe;
        const float dL_ddet0 = dL_dsqrtcoef * inv_det1e;
        const float dL_ddet1 = dL_dsqrtcoef * det_0 * (-1.f / (det_1 * det_1 + 1e-6f));
        const float dcoef_da = dL_ddet0 * cov2D.m[1][1] + dL_ddet1 * (cov2D.m[1][1] + kernel_size);
        const float dcoef_db = dL_ddet0 * (-2.f * cov2D.m[0][1]) + dL_ddet1 * (-2.f * cov2D.m[0][1]);
        const float dcoef_dc = dL_ddet0 * cov2D.m[0][0] + dL_ddet1 * (cov2D.m[0][0] + kernel_size);

        const float a = cov2D.m[0][0] += kernel_size;
        const float b = cov2D.m[0][1];
        const float c = cov2D.m[1][1] += kernel_size;

        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

            if (det_0 <= 1e-6f || det_1 <= 1e-6f) {
                dL_dopacity = 0;
            } else {
                dL_da += dcoef_da;
                dL_dc += dcoef_dc;
                dL_db += dcoef_db;
                dL_dopacity = dL_dopacity * coef;
            }

            o_cov3D[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
            o_cov3D[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
            o_cov3D[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
            o_cov3D[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
            o_cov3D[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
            o_cov3D[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
        }
        o_opacity = dL_dopacity;

        const float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
                              (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
        const float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
                              (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;

        // W (column-major): W[c][r] = view[4*r + c]
        const float W00 = s_view[0], W01 = s_view[4], W02 = s_view[8];
        const float W10 = s_view[1], W11 = s_view[5], W12 = s_view[9];
        const float W20 = s_view[2], W21 = s_view[6], W22 = s_view[10];
        const float dL_dJ00 = W00 * dL_dT00 + W01 * dL_dT01 + W02 * dL_dT02;
        const float dL_dJ02 = W20 * dL_dT00 + W21 * dL_dT01 + W22 * dL_dT02;
        const float dL_dJ11 = W10 * dL_dT10 + W11 * dL_dT11 + W12 * dL_dT12;
        const float dL_dJ12 = W20 * dL_dT10 + W21 * dL_dT11 + W22 * dL_dT12;

        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;

        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;

        // transformVec4x3Transpose (auxiliary.h:89-97)
        float3 dL_dmean_cov = {
            s_view[0] * dL_dtx + s_view[1] * dL_dty + s_view[2] * dL_dtz,
            s_view[4] * dL_dtx + s_view[5] * dL_dty + s_view[6] * dL_dtz,
            s_view[8] * dL_dtx + s_view[9] * dL_dty + s_view[10] * dL_dtz,
        };

        // ---- projective-divide backward of the screen-space mean (backward.cu:402-423) ----
        const float* proj = s_proj;
        const float3 m = mean;
        const float4 m_hom = xform_point_4x4(m, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
        float3 dL_dmean;
        dL_dmean.x = (proj[0] * m_w - proj[3] * mul1) * o_mean2D[0] + (proj[1] * m_w - proj[3] * mul2) * o_mean2D[1];
        dL_dmean.y = (proj[4] * m_w - proj[7] * mul1) * o_mean2D[0] + (proj[5] * m_w - proj[7] * mul2) * o_mean2D[1];
        dL_dmean.z = (proj[8] * m_w - proj[11] * mul1) * o_mean2D[0] + (proj[9] * m_w - proj[11] * mul2) * o_mean2D[1];

        o_mean3D[0] = dL_dmean_cov.x + dL_dmean.x;
        o_mean3D[1] = dL_dmean_cov.y + dL_dmean.y;
        o_mean3D[2] = dL_dmean_cov.z + dL_dmean.z;

        // ---- SH backward (backward.cu:20-139) ----
        if (HAS_SH) {
            const float3 g = sh_backward(p.D, p.M, p.shs + (size_t)idx * p.M * 3, p.sh_vec != 0, mean, p.campos, p.clamped[idx],
                                         V3{o_color[0], o_color[1], o_color[2]}, p.dL_dsh + (size_t)idx * p.M * 3);
            o_mean3D[0] += g.x;
            o_mean3D[1] += g.y;
            o_mean3D[2] += g.z;
            sh_written = true;
        }

        // ---- cov3D -> scale / rotation backward (backward.cu:314-377) ----
        if (p.scales != nullptr) {
            const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
            const Mat3 R = quat_to_mat(rot);
            Mat3 S = mat3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
            const float3 s = {p.scale_modifier * sc.x, p.scale_modifier * sc.y, p.scale_modifier * sc.z};
            S.m[0][0] = s.x;
            S.m[1][1] = s.y;
            S.m[2][2] = s.z;
            const Mat3 M = mat3_mul(S, R);
            const float* g6 = o_cov3D;
            const Mat3 dL_dSigma = mat3_cols(
                g6[0], 0.5f * g6[1], 0.5f * g6[2],
                0.5f * g6[1], g6[3], 0.5f * g6[4],
                0.5f * g6[2], 0.5f * g6[4], g6[5]);
            // dL_dM = 2 * M * dL_dSigma  (scalar * matrix first, as GLM evaluates it)
            Mat3 M2;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) M2.m[cc][rr] = 2.0f * M.m[cc][rr];
            const Mat3 dL_dM = mat3_mul(M2, dL_dSigma);
            const Mat3 Rt = mat3_transpose(R);
            Mat3 dL_dMt = mat3_transpose(dL_dM);

            o_scale[0] = Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2];
            o_scale[1] = Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2];
            o_scale[2] = Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2];

#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dL_dMt.m[0][k] *= s.x;
                dL_dMt.m[1][k] *= s.y;
                dL_dMt.m[2][k] *= s.z;
            }

            o_rot[0] = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
            o_rot[1] = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
            o_rot[2] = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
            o_rot[3] = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
        }
    }

    // ---- write every output row ----
    const size_t i = (size_t)idx;
    p.dL_dmean2D[3 * i + 0] = o_mean2D[0];
    p.dL_dmean2D[3 * i + 1] = o_mean2D[1];
    p.dL_dmean2D[3 * i + 2] = o_mean2D[2];
    if (p.dL_dconic) reinterpret_cast<float4*>(p.dL_dconic)[i] = make_float4(o_conic[0], o_conic[1], 0.f, o_conic[3]);
    p.dL_dopacity[i] = o_opacity;
    p.dL_dcolor[3 * i + 0] = o_color[0];
    p.dL_dcolor[3 * i + 1] = o_color[1];
    p.dL_dcolor[3 * i + 2] = o_color[2];
    p.dL_dmean3D[3 * i + 0] = o_mean3D[0];
    p.dL_dmean3D[3 * i + 1] = o_mean3D[1];
    p.dL_dmean3D[3 * i + 2] = o_mean3D[2];
    if (p.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) p.dL_dcov3D[6 * i + k] = o_cov3D[k];
    }
    if (HAS_SH && !sh_written) {
        float* row = p.dL_dsh + i * p.M * 3;
        if (p.sh_vec) {
            float4* r4 = reinterpret_cast<float4*>(row);
            for (int k = 0; k < (p.M * 3) >> 2; ++k) r4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int k = 0; k < p.M * 3; ++k) row[k] = 0.f;
        }
    }
    if (p.dL_dscale) {
        p.dL_dscale[3 * i + 0] = o_scale[0];
        p.dL_dscale[3 * i + 1] = o_scale[1];
        p.dL_dscale[3 * i + 2] = o_scale[2];
    }
    if (p.dL_drot) reinterpret_cast<float4*>(p.dL_drot)[i] = make_float4(o_rot[0], o_rot[1], o_rot[2], o_rot[3]);
}

int launch_preprocess_bwd(const GsrBackwardArgs& a, const GeomState& g, BwdAccum* accum, cudaStream_t s) {
    PreBwdParams p;
    p.P = a.P; p.D = a.D; p.M = a.M;
    p.focal_y = a.H / (2.0f * a.tan_fovy);
    p.focal_x = a.W / (2.0f * a.tan_fovx);
    p.tan_fovx = a.tan_fovx; p.tan_fovy = a.tan_fovy;
    p.kernel_size = a.kernel_size; p.scale_modifier = a.scale_modifier;
    p.means3D = a.means3D; p.radii = a.radii; p.shs = a.shs; p.clamped = g.clamped;
    p.scales = a.scales; p.rotations = a.rotations; p.cov3D_precomp = a.cov3D_precomp;
    p.view = a.viewmatrix; p.proj = a.projmatrix; p.campos = a.campos;
    p.rec = g.rec; p.accum = reinterpret_cast<const float*>(accum);
    p.dL_dmean2D = a.dL_dmean2D; p.dL_dconic = a.dL_dconic; p.dL_dopacity = a.dL_dopacity;
    p.dL_dcolor = a.dL_dcolor; p.dL_dmean3D = a.dL_dmean3D; p.dL_dcov3D = a.dL_dcov3D;
    p.dL_dsh = a.dL_dsh; p.dL_dscale = a.dL_dscale; p.dL_drot = a.dL_drot;
    static int occ = -1;
    if (occ < 0) {
        const char* e = getenv("GSR_PREBWD_OCC");    // tuning aid: resident CTAs per SM the register allocation targets
        occ = e ? atoi(e) : 4;
    }
    p.sh_vec = a.shs != nullptr && (a.M == 4 || a.M == 16) &&
               (reinterpret_cast<uintptr_t>(a.shs) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.dL_dsh) & 15u) == 0;
    if (a.shs != nullptr) preprocess_bwd_kernel<2, true><<<(a.P + 255) / 256, 256, 0, s>>>(p);   // 48 + 16 extra live registers
    else if (occ >= 4) preprocess_bwd_kernel<4, false><<<(a.P + 255) / 256, 256, 0, s>>>(p);
    else if (occ == 3) preprocess_bwd_kernel<3, false><<<(a.P + 255) / 256, 256, 0, s>>>(p);
    else preprocess_bwd_kernel<2, false><<<(a.P + 255) / 256, 256, 0, s>>>(p);
    count_launches(1);
    return 0;
}

}  // namespace gsr
