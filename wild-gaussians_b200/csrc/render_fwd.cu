// render_fwd.cu -- per-tile front-to-back alpha compositing.
// Replaces renderCUDA<3> of the reference (forward.cu:273-395).
//
// One CTA per 16x16 tile; a thread owns PPT pixels, a warp a compact pixel block (render_common.cuh).
// The tile's slice of the sorted instance list is consumed in rounds of 256 instances; each round's
// 32-byte projected records and colours are gathered into shared memory with asynchronous copies
// (cp.async / LDGSTS, no register staging) into a double buffer, two rounds of Gaussian ids ahead, so
// the gather of round r+1 overlaps the blending of round r.  Colours are staged too (the reference
// fetches them from global memory inside the blend loop, forward.cu:376).
//
// Per-warp culling: before blending a round, every lane tests 8 of the 256 staged Gaussians against the
// bounding box of the warp's sample positions using the conservative footprint extents computed in
// preprocess_fwd.cu; eight ballots give the warp a 256-bit mask and only the set bits are blended.
// The skipped (pixel, Gaussian) pairs are pairs the reference rejects with alpha < 1/255
// (forward.cu:365), so the image, final_T and n_contrib are unchanged; `contributor` is derived from
// the position in the list, not counted, so it still counts every instance like forward.cu:349.
//
// Numerics contract (SURVEY.md 8a note N3): the expression shapes of power / alpha / test_T / the
// colour accumulation are exactly the reference's (forward.cu:353-378,393), and expf is the accurate
// one, so the three discontinuous tests (power > 0, alpha < 1/255, T(1-alpha) < 1e-4) take the same
// branch and n_contrib / pixels match bit for bit.
#include "render_common.cuh"
#include <cstdlib>

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

template <int PPT, bool FLAT>
__global__ void __launch_bounds__(PixelMap<PPT>::THREADS) render_fwd_kernel(const __grid_constant__ RenderFwdParams p) {
    using PM = PixelMap<PPT>;
    constexpr int THREADS = PM::THREADS;
    __shared__ __align__(16) float4 s_geo[2][RT_BATCH];   // {x, y, hx, hy}
    __shared__ __align__(16) float4 s_con[2][RT_BATCH];   // {conic.x, conic.y, conic.z, opacity}
    __shared__ __align__(16) float4 s_col[2][RT_BATCH];   // {r, g, b, -}

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + p.ty0;

    float2 pixf[PPT];
    unsigned pix_id[PPT];
    bool inside[PPT];
    unsigned live = 0;          // bit k set: pixel k still compositing
    float bx0 = 3.0e38f, bx1 = -3.0e38f, by0 = 3.0e38f, by1 = -3.0e38f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int lx, ly;
        PM::pixel(tid, k, lx, ly);
        const unsigned px = tile_x * TILE + lx, py = tile_y * TILE + ly;
        pix_id[k] = p.W * py + px;
        inside[k] = px < (unsigned)p.W && py < (unsigned)p.H;
        pixf[k] = {(float)px, (float)py};
        if (inside[k]) {
            const float2 so = p.subpixel_offset[pix_id[k]];
            pixf[k].x += so.x;
            pixf[k].y += so.y;
            live |= 1u << k;
            bx0 = fminf(bx0, pixf[k].x); bx1 = fmaxf(bx1, pixf[k].x);
            by0 = fminf(by0, pixf[k].y); by1 = fmaxf(by1, pixf[k].y);
        }
    }
    const WarpBox box = warp_box_reduce(bx0, bx1, by0, by1);

    const uint2 range = p.ranges[tile_y * p.grid_x + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + RT_BATCH - 1) / RT_BATCH;

    // gather of one round into buffer `buf`; ids were loaded one iteration earlier
    auto stage = [&](int buf, int round, const uint32_t (&ids)[PPT]) {
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int slot = q * THREADS + tid;
            if (round * RT_BATCH + slot < total) {
                const float4* src = p.rec + 2 * (size_t)ids[q];
                cp_async_16(&s_geo[buf][slot], src);
                cp_async_16(&s_con[buf][slot], src + 1);
                const float* c = p.colors + 3 * (size_t)ids[q];
                float* cdst = reinterpret_cast<float*>(&s_col[buf][slot]);
                cp_async_4(cdst, c);
                cp_async_4(cdst + 1, c + 1);
                cp_async_4(cdst + 2, c + 2);
            }
        }
    };
    auto load_ids = [&](int round, uint32_t (&ids)[PPT]) {
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int i = round * RT_BATCH + q * THREADS + tid;
            ids[q] = (round < rounds && i < total) ? p.point_list[range.x + i] : 0u;
        }
    };

    float T[PPT], C0[PPT], C1[PPT], C2[PPT];
    uint32_t last_contributor[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) { T[k] = 1.0f; C0[k] = C1[k] = C2[k] = 0.f; last_contributor[k] = 0; }

    if (rounds > 0) {
        uint32_t ids[PPT];
        load_ids(0, ids);
        stage(0, 0, ids);
        cp_async_commit();
        load_ids(1, ids);

        int toDo = total;
        for (int r = 0; r < rounds; ++r, toDo -= RT_BATCH) {
            const int buf = r & 1;
            // this round's data has landed (for this thread) ...
            cp_async_wait<0>();
            // ... and for everyone; also the block-wide early-out vote (forward.cu:330-332)
            const int num_done = __syncthreads_count(live == 0);
            if (num_done == THREADS) break;
            // next round's gather overlaps this round's blending
            if (r + 1 < rounds) {
                stage(buf ^ 1, r + 1, ids);
                cp_async_commit();
                load_ids(r + 2, ids);
            }

            const int n = min(RT_BATCH, toDo);
            const uint32_t round_base = (uint32_t)(r * RT_BATCH);
            if (__any_sync(0xFFFFFFFFu, live != 0)) {
                // which of the staged Gaussians can touch this warp's pixels?  32 at a time: one ballot, then only
                // the set bits are blended
#pragma unroll 1
                for (int w = 0; w < RT_BATCH / 32; ++w) {
                    const int jl = w * 32 + lane;
                    unsigned mm = __ballot_sync(0xFFFFFFFFu, jl < n && box_may_touch(s_geo[buf][jl], box));
                    while (mm) {
                        const int j = w * 32 + __ffs(mm) - 1;
                        mm &= mm - 1;
                        const float4 geo = s_geo[buf][j];
                        const float4 con_o = s_con[buf][j];
                        if (FLAT) {
                            // branch-light form: every lane evaluates the pair, the state update is predicated.
                            // Same expressions, same tests, same results as the branchy form below.
#pragma unroll
                            for (int k = 0; k < PPT; ++k) {
                                const float2 d = {geo.x - pixf[k].x, geo.y - pixf[k].y};
                                const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                                const float alpha = min(0.99f, con_o.w * expf(power));
                                const float test_T = T[k] * (1 - alpha);
                                const bool blend = (live & (1u << k)) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                                const bool stop = blend && test_T < 0.0001f;
                                if (stop) live &= ~(1u << k);
                                if (blend && !stop) {
                                    const float4 col = s_col[buf][j];
                                    C0[k] += col.x * alpha * T[k];
                                    C1[k] += col.y * alpha * T[k];
                                    C2[k] += col.z * alpha * T[k];
                                    T[k] = test_T;
                                    last_contributor[k] = round_base + (uint32_t)j + 1u;
                                }
                            }
                            continue;
                        }
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            if (!(live & (1u << k))) continue;
                            const float2 xy = {geo.x, geo.y};
                            const float2 d = {xy.x - pixf[k].x, xy.y - pixf[k].y};
                            const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                            if (power > 0.0f) continue;

                            const float alpha = min(0.99f, con_o.w * expf(power));
                            if (alpha < 1.0f / 255.0f) continue;
                            const float test_T = T[k] * (1 - alpha);
                            if (test_T < 0.0001f) {
                                live &= ~(1u << k);
                                continue;
                            }
                            const float4 col = s_col[buf][j];
                            C0[k] += col.x * alpha * T[k];
                            C1[k] += col.y * alpha * T[k];
                            C2[k] += col.z * alpha * T[k];
                            T[k] = test_T;
                            // 1-based position of this instance in the tile's list (forward.cu:349,382)
                            last_contributor[k] = round_base + (uint32_t)j + 1u;
                        }
                    }
                    if (!__any_sync(0xFFFFFFFFu, live != 0)) break;
                }
            }
            // everyone is finished with `buf` before round r+2 is staged into it
            __syncthreads();
        }
        cp_async_wait<0>();
    }

    const size_t plane = (size_t)p.H * p.W;
    const float bg0 = p.bg[0], bg1 = p.bg[1], bg2 = p.bg[2];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (inside[k]) {
            p.final_T[pix_id[k]] = T[k];
            p.n_contrib[pix_id[k]] = last_contributor[k];
            p.out_color[0 * plane + pix_id[k]] = C0[k] + T[k] * bg0;
            p.out_color[1 * plane + pix_id[k]] = C1[k] + T[k] * bg1;
            p.out_color[2 * plane + pix_id[k]] = C2[k] + T[k] * bg2;
        }
    }
}

#ifndef GSR_FWD_PPT
#define GSR_FWD_PPT 1     // default pixels per thread; GSR_FWD_PPT in the environment overrides (tuning aid)
#endif

static int fwd_ppt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GSR_FWD_PPT");
        v = e ? atoi(e) : GSR_FWD_PPT;
        if (v != 1 && v != 2 && v != 4) v = GSR_FWD_PPT;
    }
    return v;
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
    static int flat = -1;
    if (flat < 0) {
        const char* e = getenv("GSR_FWD_FLAT");      // tuning aid: 1 = predicated (branch-light) blend loop
        flat = e ? atoi(e) : 0;
    }
    switch (fwd_ppt() + (flat ? 8 : 0)) {
        case 1: render_fwd_kernel<1, false><<<grid, PixelMap<1>::THREADS, 0, s>>>(p); break;
        case 4: render_fwd_kernel<4, false><<<grid, PixelMap<4>::THREADS, 0, s>>>(p); break;
        case 9: render_fwd_kernel<1, true><<<grid, PixelMap<1>::THREADS, 0, s>>>(p); break;
        case 10: render_fwd_kernel<2, true><<<grid, PixelMap<2>::THREADS, 0, s>>>(p); break;
        default: render_fwd_kernel<2, false><<<grid, PixelMap<2>::THREADS, 0, s>>>(p); break;
    }
    count_launches(1);
    return 0;
}

}  // namespace gsr
