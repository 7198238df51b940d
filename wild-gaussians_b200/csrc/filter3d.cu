// filter3d.cu -- the Mip-Splatting 3D filter size of every Gaussian over ALL training cameras in two launches
// (`compute_3D_filter`, wildgaussians/method.py:1140-1190).
//
// What it replaces: a Python loop over the C training cameras (method.py:1147-1181), per camera ~20 PyTorch launches over
// all P Gaussians (a [P,3] x [3,3] matmul, clamps, divides, four comparisons, two boolean-indexed read-modify-writes with a
// nonzero() + host synchronisation each) plus two H2D copies of the camera's R / T -- called after every densification and
// every 100 iterations late in training (method.py:2007,2012-2015).  Here: one kernel keeps a Gaussian's running minimum
// depth in a register while it walks the cameras (20 floats per camera, staged in shared memory in batches), a second tiny
// kernel applies `distance[~valid] = distance[valid].max()` and the final scale.  12 B read + 4 B written per Gaussian
// instead of ~100 B x C; the work is C x P fused-multiply-adds (issue-bound for large C, HBM-bound for small C).
//
// Arithmetic per (Gaussian, camera), in the reference's order and rounding (every PyTorch statement is its own kernel,
// so there is no contraction ACROSS statements; inside the matmul cuBLAS accumulates x R0j, y R1j, z R2j left to right):
//     cam   = xyz @ R + T                      c_j = fma(z, R2j, fma(y, R1j, x * R0j)) + T_j
//     valid = cam.z > 0.2  and  -0.15 W <= px <= 1.15 W  and  -0.15 H <= py <= 1.15 H
//             with zc = max(cam.z, 0.001), px = cam.x / zc * fx + W / 2, py = cam.y / zc * fy + H / 2
//     distance = min(distance, zc) where valid                      (distance starts at 100000)
// then  distance[~valid_any] = max(distance[valid_any]);  filter_3D = distance / max_c(fx) * sqrt(0.2).
#include "common.cuh"
#include <cfloat>
#include <cmath>

namespace gsr {

constexpr int F3_THREADS = 256;
constexpr int F3_CAM_FLOATS = 20;          // per camera: R (9, row-major), T (3), fx, fy, W/2, H/2, x_lo, x_hi, y_lo, y_hi
constexpr int F3_CAM_BATCH = 256;          // cameras staged per round: 256 x 80 B = 20 KB

// dmax_bits: max over the Gaussians seen by at least one camera of their distance, as the bit pattern of a positive float
// (unsigned order = float order); 0 when no Gaussian is seen by any camera
__global__ void __launch_bounds__(F3_THREADS) filter3d_distance_kernel(int P, const float* __restrict__ xyz, int C,
                                                                       const float* __restrict__ cams, float* __restrict__ distance,
                                                                       unsigned* __restrict__ dmax_bits) {
    __shared__ __align__(16) float s_cam[F3_CAM_BATCH][F3_CAM_FLOATS];
    __shared__ unsigned s_max;
    const int i = blockIdx.x * F3_THREADS + threadIdx.x;
    if (threadIdx.x == 0) s_max = 0u;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < P) { x = xyz[3 * (size_t)i]; y = xyz[3 * (size_t)i + 1]; z = xyz[3 * (size_t)i + 2]; }
    float dist = 100000.0f;
    bool seen = false;
    for (int c0 = 0; c0 < C; c0 += F3_CAM_BATCH) {
        const int nb = min(F3_CAM_BATCH, C - c0);
        __syncthreads();
        for (int k = threadIdx.x; k < nb * (F3_CAM_FLOATS / 4); k += F3_THREADS)
            reinterpret_cast<float4*>(&s_cam[0][0])[k] = reinterpret_cast<const float4*>(cams + (size_t)c0 * F3_CAM_FLOATS)[k];
        __syncthreads();
        if (i < P) {
#pragma unroll 4
            for (int c = 0; c < nb; ++c) {
                const float4 r0 = *reinterpret_cast<const float4*>(&s_cam[c][0]);    // R00 R01 R02 R10
                const float4 r1 = *reinterpret_cast<const float4*>(&s_cam[c][4]);    // R11 R12 R20 R21
                const float4 r2 = *reinterpret_cast<const float4*>(&s_cam[c][8]);    // R22 T0 T1 T2
                const float4 k4 = *reinterpret_cast<const float4*>(&s_cam[c][12]);   // fx fy W/2 H/2
                const float cz = __fadd_rn(fmaf(z, r2.x, fmaf(y, r1.y, __fmul_rn(x, r0.z))), r2.w);
                if (!(cz > 0.2f)) continue;
                const float cx = __fadd_rn(fmaf(z, r1.z, fmaf(y, r0.w, __fmul_rn(x, r0.x))), r2.y);
                const float cy = __fadd_rn(fmaf(z, r1.w, fmaf(y, r1.x, __fmul_rn(x, r0.y))), r2.z);
                const float zc = fmaxf(cz, 0.001f);
                const float px = __fadd_rn(__fmul_rn(__fdiv_rn(cx, zc), k4.x), k4.z);
                const float py = __fadd_rn(__fmul_rn(__fdiv_rn(cy, zc), k4.y), k4.w);
                // bounds -0.15 W, 1.15 W, -0.15 H, 1.15 H: fp32 roundings of the host's double products (what the comparisons
                // of a float32 tensor with a Python scalar use)
                const float4 bd = *reinterpret_cast<const float4*>(&s_cam[c][16]);
                if (px >= bd.x && px <= bd.y && py >= bd.z && py <= bd.w) {
                    dist = fminf(dist, zc);
                    seen = true;
                }
            }
        }
    }
    if (i < P) {
        // provisional: Gaussians no camera sees are marked with a negative distance for the second kernel
        distance[i] = seen ? dist : -1.0f;
    }
    const unsigned bits = (i < P && seen) ? __float_as_uint(dist) : 0u;
    const unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, bits);
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(&s_max, wmax);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(dmax_bits, s_max);
}

__global__ void __launch_bounds__(F3_THREADS) filter3d_finish_kernel(int P, const float* __restrict__ distance,
                                                                     const unsigned* __restrict__ dmax_bits, float focal_length,
                                                                     float scale, float* __restrict__ filter_3D) {
    const int i = blockIdx.x * F3_THREADS + threadIdx.x;
    if (i >= P) return;
    float d = distance[i];
    if (d < 0.f) d = __uint_as_float(*dmax_bits);            // distance[~valid_points] = distance[valid_points].max()
    filter_3D[i] = __fmul_rn(__fdiv_rn(d, focal_length), scale);     // distance / focal_length * (0.2 ** 0.5)
}

}  // namespace gsr

extern "C" size_t gsr_filter3d_scratch_bytes(int P) { return (size_t)(P > 0 ? P : 0) * sizeof(float) + 256; }

extern "C" int gsr_compute_3d_filter(int P, const float* xyz, int num_cameras, const float* cameras, float focal_length,
                                     float* filter_3D, void* scratch, void* stream) {
    using namespace gsr;
    if (P < 0 || num_cameras < 0) { set_error("bad P / num_cameras"); return GSR_E_INVALID; }
    if (P == 0) return 0;
    if (!xyz || !filter_3D || !scratch || (num_cameras > 0 && !cameras)) { set_error("a required pointer is NULL"); return GSR_E_INVALID; }
    if ((reinterpret_cast<uintptr_t>(cameras) | reinterpret_cast<uintptr_t>(scratch)) & 15u) { set_error("cameras / scratch must be 16-byte aligned"); return GSR_E_INVALID; }
    if (!(focal_length > 0.f)) { set_error("focal_length must be positive"); return GSR_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* dmax = reinterpret_cast<unsigned*>(scratch);
    float* distance = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 256);
    GSR_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned), s));
    const int grid = (P + F3_THREADS - 1) / F3_THREADS;
    filter3d_distance_kernel<<<grid, F3_THREADS, 0, s>>>(P, xyz, num_cameras, cameras, distance, dmax);
    filter3d_finish_kernel<<<grid, F3_THREADS, 0, s>>>(P, distance, dmax, focal_length, (float)sqrt(0.2), filter_3D);   // float32(0.2 ** 0.5)
    count_launches(2);
    GSR_CUDA(cudaGetLastError());
    return 0;
}
