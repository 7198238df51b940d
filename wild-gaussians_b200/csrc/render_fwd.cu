// render_fwd.cu -- per-tile front-to-back alpha compositing.
// Replaces renderCUDA<3> of the reference (forward.cu:273-395).
//
// One CTA per 16x16 tile, one thread per pixel, each warp owning a compact 8x4 pixel block.  The tile's
// slice of the sorted instance list is consumed in rounds of 256 instances; each round's 32-byte
// projected records and colours are gathered into shared memory with asynchronous copies
// (cp.async / LDGSTS, no register staging) into a double buffer, two rounds of Gaussian ids ahead, so
// the gather of round r+1 overlaps the blending of round r.  Colours are staged too (the reference
// fetches them from global memory inside the blend loop, forward.cu:376).
//
// Per-warp culling: before blending a round, every lane tests 8 of the 256 staged Gaussians against the
// bounding box of the warp's 32 sample positions using the conservative footprint extents computed
// in preprocess_fwd.cu; eight ballots give the warp a 256-bit mask and only the set bits are blended.
// The skipped (pixel, Gaussian) pairs are pairs the reference rejects with alpha < 1/255
// (forward.cu:365), so the image, final_T and n_contrib are unchanged; `contributor` is derived from
// the position in the list, not counted, so it still counts every instance like forward.cu:349.
//
// Numerics contract (SURVEY.md 8a note N3): the expression shapes of power / alpha / test_T / the
// colour accumulation are exactly the reference's (forward.cu:353-378,393), and expf is the accurate
// one, so the three discontinuous tests (power > 0, alpha < 1/255, T(1-alpha) < 1e-4) take the same
// branch and n_contrib / pixels match bit for bit.
#include "render_common.cuh"

namespace gsr {

struct RenderFwdParams {
    int W, H, grid_x, ty0;
    const uint2* ranges;
    const uint32_t* point_list;
    const float2* subpixel_offset;
    const float4* rec;
    const float* colors;       // [P,3]
    const float* bg;
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
};

__global__ void __launch_bounds__(RT_THREADS) render_fwd_kernel(const __grid_constant__ RenderFwdParams p) {
    __shared__ __align__(16) float4 s_geo[2][RT_THREADS];   // {x, y, hx, hy}
    __shared__ __align__(16) float4 s_con[2][RT_THREADS];   // {conic.x, conic.y, conic.z, opacity}
    __shared__ float s_col[2][RT_THREADS][3];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + p.ty0;
    int lx, ly;
    tile_pixel(tid, lx, ly);
    const unsigned pix_x = tile_x * TILE + lx;
    const unsigned pix_y = tile_y * TILE + ly;
    const unsigned pix_id = p.W * pix_y + pix_x;
    const bool inside = pix_x < (unsigned)p.W && pix_y < (unsigned)p.H;
    bool done = !inside;

    float2 pixf = {(float)pix_x, (float)pix_y};
    if (inside) {
        const float2 so = p.subpixel_offset[pix_id];
        pixf.x += so.x;
        pixf.y += so.y;
    }
    const WarpBox box = warp_box(pixf, inside);

    const uint2 range = p.ranges[tile_y * p.grid_x + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + RT_THREADS - 1) / RT_THREADS;

    // gather of one round into buffer `buf`; `id` was loaded one iteration earlier
    auto stage = [&](int buf, int round, uint32_t id) {
        if (round * RT_THREADS + tid < total) {
            const float4* src = p.rec + 2 * (size_t)id;
            cp_async_16(&s_geo[buf][tid], src);
            cp_async_16(&s_con[buf][tid], src + 1);
            const float* c = p.colors + 3 * (size_t)id;
            cp_async_4(&s_col[buf][tid][0], c);
            cp_async_4(&s_col[buf][tid][1], c + 1);
            cp_async_4(&s_col[buf][tid][2], c + 2);
        }
    };
    auto load_id = [&](int round) -> uint32_t {
        const int i = round * RT_THREADS + tid;
        return (round < rounds && i < total) ? p.point_list[range.x + i] : 0u;
    };

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;

    if (rounds > 0) {
        uint32_t id_next = load_id(0);
        stage(0, 0, id_next);
        cp_async_commit();
        id_next = load_id(1);

        int toDo = total;
        for (int r = 0; r < rounds; ++r, toDo -= RT_THREADS) {
            const int buf = r & 1;
            // this round's data has landed (for this thread) ...
            cp_async_wait<0>();
            // ... and for everyone; also the block-wide early-out vote (forward.cu:330-332)
            const int num_done = __syncthreads_count(done);
            if (num_done == RT_THREADS) break;
            // next round's gather overlaps this round's blending
            if (r + 1 < rounds) {
                stage(buf ^ 1, r + 1, id_next);
                cp_async_commit();
                id_next = load_id(r + 2);
            }

            const int n = min(RT_THREADS, toDo);
            const uint32_t round_base = (uint32_t)(r * RT_THREADS);
            if (!__all_sync(0xFFFFFFFFu, done)) {
                // which of the staged Gaussians can touch this warp's pixels?
                unsigned mask[RT_THREADS / 32];
#pragma unroll
                for (int w = 0; w < RT_THREADS / 32; ++w) {
                    const int j = w * 32 + lane;
                    mask[w] = __ballot_sync(0xFFFFFFFFu, j < n && box_may_touch(s_geo[buf][j], box));
                }
#pragma unroll
                for (int w = 0; w < RT_THREADS / 32; ++w) {
                    unsigned mm = mask[w];
                    while (mm) {
                        const int j = w * 32 + __ffs(mm) - 1;
                        mm &= mm - 1;
                        if (done) continue;
                        const float4 geo = s_geo[buf][j];
                        const float2 xy = {geo.x, geo.y};
                        const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                        const float4 con_o = s_con[buf][j];
                        const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                        if (power > 0.0f) continue;

                        const float alpha = min(0.99f, con_o.w * expf(power));
                        if (alpha < 1.0f / 255.0f) continue;
                        const float test_T = T * (1 - alpha);
                        if (test_T < 0.0001f) {
                            done = true;
                            continue;
                        }
                        C0 += s_col[buf][j][0] * alpha * T;
                        C1 += s_col[buf][j][1] * alpha * T;
                        C2 += s_col[buf][j][2] * alpha * T;
                        T = test_T;
                        // 1-based position of this instance in the tile's list (forward.cu:349,382)
                        last_contributor = round_base + (uint32_t)j + 1u;
                    }
                    if (__all_sync(0xFFFFFFFFu, done)) break;
                }
            }
            // everyone is finished with `buf` before round r+2 is staged into it
            __syncthreads();
        }
        cp_async_wait<0>();
    }

    if (inside) {
        p.final_T[pix_id] = T;
        p.n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)p.H * p.W;
        p.out_color[0 * plane + pix_id] = C0 + T * p.bg[0];
        p.out_color[1 * plane + pix_id] = C1 + T * p.bg[1];
        p.out_color[2 * plane + pix_id] = C2 + T * p.bg[2];
    }
}

int launch_render_fwd(const GsrForwardArgs& a, const GeomState& g, const BinState& b, const ImgState& im,
                      const float* colors, int ty0, int ty1, cudaStream_t s) {
    RenderFwdParams p;
    p.W = a.W; p.H = a.H; p.grid_x = tiles_x(a.W); p.ty0 = ty0;
    p.ranges = im.ranges; p.point_list = b.point_list;
    p.subpixel_offset = reinterpret_cast<const float2*>(a.subpixel_offset);
    p.rec = g.rec; p.colors = colors; p.bg = a.background;
    p.final_T = im.final_T; p.n_contrib = im.n_contrib; p.out_color = a.out_color;
    if (ty1 <= ty0) return 0;
    dim3 grid(p.grid_x, ty1 - ty0, 1);
    render_fwd_kernel<<<grid, RT_THREADS, 0, s>>>(p);
    count_launches(1);
    return 0;
}

}  // namespace gsr
