// render_common.cuh -- pieces shared by the forward and backward compositing kernels.
#pragma once
#include "common.cuh"

namespace gsr {

constexpr int RT_THREADS = 256;          // one thread per pixel of a 16x16 tile
constexpr int RT_WARPS = RT_THREADS / 32;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Thread -> pixel mapping inside a tile: each warp owns a compact 8x4 pixel block (2 x 4 blocks per
// tile) instead of the reference's 16x2 strip.  Per-pixel results do not depend on the mapping; the
// compact block makes the per-warp culling below reject more Gaussians and keeps every 32-byte output
// sector written by one warp.
__device__ __forceinline__ void tile_pixel(int tid, int& lx, int& ly) {
    const int warp = tid >> 5, lane = tid & 31;
    lx = (warp & 1) * 8 + (lane & 7);
    ly = (warp >> 1) * 4 + (lane >> 3);
}

// Bounding box of the warp's (sub-pixel shifted) sample positions; lanes outside the image are ignored.
struct WarpBox {
    float x0, x1, y0, y1;
};
__device__ __forceinline__ WarpBox warp_box(float2 pixf, bool inside) {
    float x0 = inside ? pixf.x : 3.0e38f, x1 = inside ? pixf.x : -3.0e38f;
    float y0 = inside ? pixf.y : 3.0e38f, y1 = inside ? pixf.y : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xFFFFFFFFu, x0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xFFFFFFFFu, x1, o));
        y0 = fminf(y0, __shfl_xor_sync(0xFFFFFFFFu, y0, o));
        y1 = fmaxf(y1, __shfl_xor_sync(0xFFFFFFFFu, y1, o));
    }
    return {x0, x1, y0, y1};
}

// Can the Gaussian with centre/extents g = {x, y, hx, hy} reach alpha >= 1/255 anywhere in the box?
// Conservative (see preprocess_fwd.cu); every comparison is false for NaN, which keeps the Gaussian.
__device__ __forceinline__ bool box_may_touch(const float4 g, const WarpBox& b) {
    return !(g.z < 0.f || g.x + g.z < b.x0 || g.x - g.z > b.x1 || g.y + g.w < b.y0 || g.y - g.w > b.y1);
}

}  // namespace gsr
