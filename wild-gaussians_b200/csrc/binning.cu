// binning.cu -- (Gaussian, tile) instance emission in depth order and per-tile range
// detection.  Replaces duplicateWithKeys (rasterizer_impl.cu:70-111) and identifyTileRanges
// (:116-138) of the reference.
//
// The reference emits one 64-bit key (tile << 32 | depth bits) + 32-bit value per instance
// in Gaussian-index order and radix-sorts all R of them over 41-48 key bits (6 passes, about
// 150 B of HBM traffic per instance).  Here the Gaussians were already ordered by depth
// (4 radix passes over P, not R, pairs), so instances are emitted front-to-back and only
// have to be stably partitioned by their 32-bit tile id (2 radix passes).  The final order
// (tile-major, depth-minor, ties by Gaussian index) is identical -- a stable sort is unique.
//
// Emission is warp-cooperative: the 32 Gaussians of a warp are expanded by all 32 lanes
// together, so a splat covering 1000 tiles costs 32 coalesced iterations instead of one
// thread's 1000-iteration loop (the load imbalance noted in SURVEY.md 8a row a9).
#include "common.cuh"

namespace gsr {

__global__ void __launch_bounds__(256) emit_instances_kernel(int P, int grid_x, const uint32_t* __restrict__ order,
                                                             const uint32_t* __restrict__ offsets,
                                                             const TileRect* __restrict__ rect,
                                                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // position in depth order
    const int lane = threadIdx.x & 31;
    uint32_t g = 0, off = 0, x0 = 0, y0 = 0, w = 0, cnt = 0;
    if (i < P) {
        g = order[i];
        const TileRect r = rect[g];
        w = (uint32_t)r.x1 - r.x0;
        cnt = w * ((uint32_t)r.y1 - r.y0);
        x0 = r.x0;
        y0 = r.y0;
        off = offsets[i];
    }
    const unsigned have = __ballot_sync(0xFFFFFFFFu, cnt != 0);
    unsigned todo = have;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t sg = __shfl_sync(0xFFFFFFFFu, g, src);
        const uint32_t so = __shfl_sync(0xFFFFFFFFu, off, src);
        const uint32_t sx = __shfl_sync(0xFFFFFFFFu, x0, src);
        const uint32_t sy = __shfl_sync(0xFFFFFFFFu, y0, src);
        const uint32_t sw = __shfl_sync(0xFFFFFFFFu, w, src);
        const uint32_t sc = __shfl_sync(0xFFFFFFFFu, cnt, src);
        // instance k of this Gaussian covers tile (sy + k / sw, sx + k % sw): row-major over
        // the rectangle, the emission order of the reference (rasterizer_impl.cu:96-108)
        for (uint32_t k = lane; k < sc; k += 32) {
            const uint32_t ry = k / sw, rx = k - ry * sw;
            keys[so + k] = (sy + ry) * (uint32_t)grid_x + (sx + rx);
            vals[so + k] = sg;
        }
    }
}

int launch_emit_instances(const GeomState& g, int P, int gx, uint32_t* keys, uint32_t* vals, cudaStream_t s) {
    if (P == 0) return 0;
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, gx, g.order, g.offsets, g.rect, keys, vals);
    count_launches(1);
    return 0;
}

// ranges[t] = [first, last+1) of tile t in the sorted instance list; empty tiles keep (0,0)
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ tile_keys, size_t R,
                                                          uint2* __restrict__ ranges) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = tile_keys[i];
    if (i == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = tile_keys[i - 1];
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

int launch_tile_ranges(const uint32_t* sorted_tile_keys, size_t R, uint2* ranges, int num_tiles, cudaStream_t s) {
    GSR_CUDA(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s));
    if (R == 0) return 0;
    tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(sorted_tile_keys, R, ranges);
    count_launches(1);
    return 0;
}

}  // namespace gsr
