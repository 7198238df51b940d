// render_bwd.cu -- per-tile back-to-front re-composite producing per-Gaussian partial
// gradients dL/d{mean2D (+abs channel), conic, opacity*coef, colour}.
// Replaces renderCUDA<3> of the reference backward (backward.cu:435-606).
//
// Same per-pixel recurrence as the reference (T reconstructed by division, running
// accum_rec / last_alpha / last_color), but the accumulation is re-designed: the reference
// issues 10 global float atomics per blended (pixel, Gaussian) pair (backward.cu:568-603).
// Here all 32 lanes of a warp walk the same Gaussian in lock-step, so the 10 per-lane terms
// are summed across the warp with a multi-value shuffle butterfly (12 shuffles instead of
// 50), then added to a per-tile shared-memory accumulator (one shared atomic per component
// per warp), and a tile writes each visited Gaussian's 10 sums to HBM once, as three
// 128-bit vector reductions (red.global.add.v4.f32) into a packed 48-byte record.
// That is up to 256x fewer L2 atomics per Gaussian per tile.
//
// The walk starts at the tile's largest n_contrib instead of the end of the tile's range:
// instances behind every pixel's last contributor are skipped by the reference one by one
// (backward.cu:531-533); skipping them wholesale gives the same sums.
//
// The |grad| channel (backward.cu:593-595) is the sum over pixels of |gx| + |gy|, so the
// absolute value is taken per lane BEFORE the warp reduction.
#include "render_common.cuh"
#include <cstdlib>

namespace gsr {

constexpr int RB_ACC = 12;   // floats per Gaussian in the shared / global accumulator

struct RenderBwdParams {
    int W, H, grid_x, ty0;
    const uint2* ranges;
    const uint32_t* point_list;
    const float2* subpixel_offset;
    const float* bg;
    const float4* rec;
    const float* colors;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* dL_dpix;
    float* accum;   // [P][12]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Sum 10 per-lane values over the 32 lanes of the warp with a multi-value butterfly: at every step a
// lane keeps half of its values and ships the other half to its partner, so the exchange costs
// 5 + 3 + 2 + 1 + 1 = 12 shuffles instead of 10 x 5.  On return lane l holds in v[0] the warp total of
// component `slot` (returned); slots >= 10 are padding; every component is held by lanes l and l^1.
__device__ __forceinline__ int butterfly10(float (&v)[10], int lane) {
    float w[5], x[3], y[2];
    {   // xor 16: components (i, i+5)
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float send = hi ? v[i] : v[i + 5];
            const float keep = hi ? v[i + 5] : v[i];
            w[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
        }
    }
    {   // xor 8: (0,3) (1,4) (2,-)
        const bool hi = lane & 8;
        x[0] = (hi ? w[3] : w[0]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? w[0] : w[3], 8);
        x[1] = (hi ? w[4] : w[1]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? w[1] : w[4], 8);
        x[2] = (hi ? 0.f : w[2]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? w[2] : 0.f, 8);
    }
    {   // xor 4: (0,2) (1,-)
        const bool hi = lane & 4;
        y[0] = (hi ? x[2] : x[0]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? x[0] : x[2], 4);
        y[1] = (hi ? 0.f : x[1]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? x[1] : 0.f, 4);
    }
    {   // xor 2: (0,1)
        const bool hi = lane & 2;
        v[0] = (hi ? y[1] : y[0]) + __shfl_xor_sync(0xFFFFFFFFu, hi ? y[0] : y[1], 2);
    }
    v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], 1);
    // which component did this lane end up with?
    //   bit4: +5; bit3: slots {0,1,2} -> {3,4,pad}; bit2: {0,1} -> {2,pad}; bit1: {0} -> {1}
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    // index within w[0..4] after step 1, through steps 2-4
    int idx;            // position in the 5-vector w
    if (!b2) idx = b1 ? 1 : 0;       // y[0] (lo of xor 4) came from x[0]; y[1] from x[1]
    else idx = b1 ? -1 : 2;          // hi of xor 4: y[0] = x[2], y[1] = padding
    if (idx >= 0 && b3) idx = idx == 2 ? -1 : idx + 3;   // hi of xor 8: x[0] = w[3], x[1] = w[4], x[2] = padding
    return idx < 0 ? 15 : idx + 5 * b4;
}

template <int PPT>
__global__ void __launch_bounds__(PixelMap<PPT>::THREADS) render_bwd_kernel(const __grid_constant__ RenderBwdParams p) {
    using PM = PixelMap<PPT>;
    constexpr int THREADS = PM::THREADS;
    constexpr int WARPS = PM::WARPS;
    __shared__ uint32_t s_id[RT_BATCH];
    __shared__ __align__(16) float4 s_geo[RT_BATCH];   // {x, y, hx, hy}
    __shared__ __align__(16) float4 s_con[RT_BATCH];   // {conic.x, conic.y, conic.z, opacity}
    __shared__ float s_col[RT_BATCH][3];
    __shared__ __align__(16) float s_acc[RT_BATCH][RB_ACC];
    __shared__ uint32_t s_max[WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + p.ty0;
    const size_t plane = (size_t)p.H * p.W;

    float2 pixf[PPT];
    float T_final[PPT], T[PPT], last_alpha[PPT];
    uint32_t last_contributor[PPT];
    float accum_rec[PPT][3], last_color[PPT][3], dL_dpixel[PPT][3];
    float bx0 = 3.0e38f, bx1 = -3.0e38f, by0 = 3.0e38f, by1 = -3.0e38f;
    uint32_t thread_last = 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int lx, ly;
        PM::pixel(tid, k, lx, ly);
        const unsigned px = tile_x * TILE + lx, py = tile_y * TILE + ly;
        const unsigned pix_id = p.W * py + px;
        const bool inside = px < (unsigned)p.W && py < (unsigned)p.H;
        pixf[k] = {(float)px, (float)py};
        T_final[k] = 0.f;
        last_contributor[k] = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { accum_rec[k][ch] = 0.f; last_color[k][ch] = 0.f; dL_dpixel[k][ch] = 0.f; }
        last_alpha[k] = 0.f;
        if (inside) {
            const float2 so = p.subpixel_offset[pix_id];
            pixf[k].x += so.x;
            pixf[k].y += so.y;
            T_final[k] = p.final_T[pix_id];
            last_contributor[k] = p.n_contrib[pix_id];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) dL_dpixel[k][ch] = p.dL_dpix[ch * plane + pix_id];
            bx0 = fminf(bx0, pixf[k].x); bx1 = fmaxf(bx1, pixf[k].x);
            by0 = fminf(by0, pixf[k].y); by1 = fmaxf(by1, pixf[k].y);
        }
        T[k] = T_final[k];
        thread_last = max(thread_last, last_contributor[k]);
    }
    const WarpBox box = warp_box_reduce(bx0, bx1, by0, by1);
    const uint2 range = p.ranges[tile_y * p.grid_x + tile_x];

    // warp-wide and tile-wide largest last_contributor: nothing behind it contributes
    const uint32_t warp_last = __reduce_max_sync(0xFFFFFFFFu, thread_last);
    if (lane == 0) s_max[warp] = warp_last;
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) tile_last = max(tile_last, s_max[w]);
    const int total = (int)min(tile_last, range.y - range.x);   // instances [range.x, range.x+total) matter
    const int rounds = (total + RT_BATCH - 1) / RT_BATCH;

    const float bg0 = p.bg[0], bg1 = p.bg[1], bg2 = p.bg[2];
    // Gradient of pixel coordinate w.r.t. normalized screen-space viewport coordinates (-1 to 1)
    const float ddelx_dx = 0.5 * p.W;
    const float ddely_dy = 0.5 * p.H;

    int toDo = total;
    for (int r = 0; r < rounds; ++r, toDo -= RT_BATCH) {
        __syncthreads();   // previous round's accumulators flushed, staging buffers free
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int slot = q * THREADS + tid;
            const int progress = r * RT_BATCH + slot;
            if (progress < total) {
                const uint32_t id = p.point_list[range.x + total - progress - 1];
                s_id[slot] = id;
                const float4* src = p.rec + 2 * (size_t)id;
                s_geo[slot] = src[0];
                s_con[slot] = src[1];
                s_col[slot][0] = p.colors[3 * (size_t)id + 0];
                s_col[slot][1] = p.colors[3 * (size_t)id + 1];
                s_col[slot][2] = p.colors[3 * (size_t)id + 2];
            }
#pragma unroll
            for (int k = 0; k < RB_ACC; ++k) s_acc[slot][k] = 0.f;
        }
        __syncthreads();

        const int n = min(RT_BATCH, toDo);
        // 0-based list position of staged entry j: pos(j) = first_pos - j  (walking back to front)
        const int first_pos = total - 1 - r * RT_BATCH;

        // per-warp culling: staged Gaussians that are in front of this warp's deepest contributor and whose
        // alpha >= 1/255 footprint can touch the warp's pixel block
        unsigned mask[RT_BATCH / 32];
#pragma unroll
        for (int w = 0; w < RT_BATCH / 32; ++w) {
            const int j = w * 32 + lane;
            const bool rel = j < n && (uint32_t)(first_pos - j) < warp_last && box_may_touch(s_geo[j], box);
            mask[w] = __ballot_sync(0xFFFFFFFFu, rel);
        }
#pragma unroll
        for (int w = 0; w < RT_BATCH / 32; ++w) {
            unsigned mm = mask[w];
            while (mm) {
                const int j = w * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t pos = (uint32_t)(first_pos - j);
                const float4 geo = s_geo[j];
                const float4 con_o = s_con[j];
                const float c0 = s_col[j][0], c1 = s_col[j][1], c2 = s_col[j][2];

                float v[10];
#pragma unroll
                for (int q = 0; q < 10; ++q) v[q] = 0.f;
                bool any_active = false;
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    // the reference visits an instance iff its position is below the pixel's n_contrib
                    // (backward.cu:529-533); pixels outside the image have n_contrib = 0
                    if (!(pos < last_contributor[k])) continue;
                    const float2 d = {geo.x - pixf[k].x, geo.y - pixf[k].y};
                    const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = min(0.99f, con_o.w * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    any_active = true;

                    // one reciprocal serves T / (1 - alpha) and T_final / (1 - alpha) (backward.cu:548,579);
                    // gradients are compared at 1e-3, the reciprocal is good to 1 ulp
                    const float inv = __frcp_rn(1.f - alpha);
                    T[k] = T[k] * inv;
                    const float dchannel_dcolor = alpha * T[k];
                    float dL_dalpha = 0.0f;
                    const float cc[3] = {c0, c1, c2};
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        accum_rec[k][ch] = last_alpha[k] * last_color[k][ch] + (1.f - last_alpha[k]) * accum_rec[k][ch];
                        last_color[k][ch] = cc[ch];
                        const float dL_dchannel = dL_dpixel[k][ch];
                        dL_dalpha += (cc[ch] - accum_rec[k][ch]) * dL_dchannel;
                        v[7 + ch] += dchannel_dcolor * dL_dchannel;
                    }
                    dL_dalpha *= T[k];
                    last_alpha[k] = alpha;

                    const float bg_dot_dpixel = bg0 * dL_dpixel[k][0] + bg1 * dL_dpixel[k][1] + bg2 * dL_dpixel[k][2];
                    dL_dalpha += (-T_final[k] * inv) * bg_dot_dpixel;

                    const float dL_dG = con_o.w * dL_dalpha;
                    const float gdx = G * d.x;
                    const float gdy = G * d.y;
                    const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
                    const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;

                    const float gx = dL_dG * dG_ddelx * ddelx_dx;
                    const float gy = dL_dG * dG_ddely * ddely_dy;
                    v[0] += gx;
                    v[1] += gy;
                    v[2] += fabsf(gx) + fabsf(gy);
                    v[3] += -0.5f * gdx * d.x * dL_dG;
                    v[4] += -0.5f * gdx * d.y * dL_dG;
                    v[5] += -0.5f * gdy * d.y * dL_dG;
                    v[6] += G * dL_dalpha;
                }
                if (!__any_sync(0xFFFFFFFFu, any_active)) continue;   // warp-uniform

                const int slot = butterfly10(v, lane);
                if ((lane & 1) == 0 && slot < 10) atomicAdd(&s_acc[j][slot], v[0]);
            }
        }
        __syncthreads();
        // one flush per visited Gaussian per tile: three 128-bit reductions (skipped when nothing landed)
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int slot = q * THREADS + tid;
            if (slot < n) {
                const float* a = s_acc[slot];
                const float4 a0 = *reinterpret_cast<const float4*>(a);
                const float4 a1 = *reinterpret_cast<const float4*>(a + 4);
                const float4 a2 = *reinterpret_cast<const float4*>(a + 8);
                const bool any = (a0.x != 0.f) | (a0.y != 0.f) | (a0.z != 0.f) | (a0.w != 0.f) | (a1.x != 0.f) |
                                 (a1.y != 0.f) | (a1.z != 0.f) | (a1.w != 0.f) | (a2.x != 0.f) | (a2.y != 0.f);
                if (any) {
                    float* dst = p.accum + (size_t)s_id[slot] * RB_ACC;
                    red_add_v4(dst + 0, a0.x, a0.y, a0.z, a0.w);
                    red_add_v4(dst + 4, a1.x, a1.y, a1.z, a1.w);
                    red_add_v4(dst + 8, a2.x, a2.y, 0.f, 0.f);
                }
            }
        }
    }
}

#ifndef GSR_BWD_PPT
#define GSR_BWD_PPT 1     // default pixels per thread; GSR_BWD_PPT in the environment overrides (tuning aid)
#endif

static int bwd_ppt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GSR_BWD_PPT");
        v = e ? atoi(e) : GSR_BWD_PPT;
        if (v != 1 && v != 2 && v != 4) v = GSR_BWD_PPT;
    }
    return v;
}

int launch_render_bwd(const GsrBackwardArgs& a, const GeomState& g, const BinState& b, const ImgState& im,
                      const float* colors, BwdAccum* accum, int ty0, int ty1, cudaStream_t s) {
    RenderBwdParams p;
    p.W = a.W; p.H = a.H; p.grid_x = tiles_x(a.W); p.ty0 = ty0;
    p.ranges = im.ranges; p.point_list = b.point_list;
    p.subpixel_offset = reinterpret_cast<const float2*>(a.subpixel_offset);
    p.bg = a.background; p.rec = g.rec; p.colors = colors;
    p.final_T = im.final_T; p.n_contrib = im.n_contrib; p.dL_dpix = a.dL_dpix;
    p.accum = reinterpret_cast<float*>(accum);
    if (ty1 <= ty0) return 0;
    dim3 grid(p.grid_x, ty1 - ty0, 1);
    switch (bwd_ppt()) {
        case 1: render_bwd_kernel<1><<<grid, PixelMap<1>::THREADS, 0, s>>>(p); break;
        case 4: render_bwd_kernel<4><<<grid, PixelMap<4>::THREADS, 0, s>>>(p); break;
        default: render_bwd_kernel<2><<<grid, PixelMap<2>::THREADS, 0, s>>>(p); break;
    }
    count_launches(1);
    return 0;
}

}  // namespace gsr
