// preprocess_fwd.cu -- per-Gaussian forward stage: near cull, 3D->2D covariance projection
// with the Mip-Splatting low-pass + opacity compensation, conic, screen radius, tile
// rectangle, optional SH->RGB.  One thread per Gaussian; HBM-bound (about 44 B in, 56 B out
// per Gaussian on the colors_precomp path).
//
// Replaces preprocessCUDA<3> of the reference (forward.cu:166-268) together with
// computeColorFromSH (forward.cu:20-71), computeCov2D (:74-124), computeCov3D (:129-163),
// in_frustum / ndc2Pix / getRect (auxiliary.h:41-56,139-164).  Differences in data layout,
// not in arithmetic: the projected state is written as one 32-byte record per Gaussian
// ({x, y, hx, hy | conic.x, conic.y, conic.z, opacity*coef}; hx, hy = conservative half extents of
// the alpha >= 1/255 footprint) so the compositing kernels gather one aligned sector per instance; cov3D is not stored (the backward recomputes it);
// the tile rectangle is stored (8 B) so instance emission does not redo getRect; the depth
// sort key is emitted here.
#include "common.cuh"
#include "gaussian_math.cuh"

namespace gsr {

struct PreFwdParams {
    int P, D, M, W, H;
    int grid_x, grid_y;
    int ty0, ty1;                 // tile-row shard
    float focal_x, focal_y, tan_fovx, tan_fovy, kernel_size, scale_modifier;
    int prefiltered;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* campos;
    const float* view;            // [16] device
    const float* proj;            // [16] device
    // outputs
    int* radii;
    float4* rec;
    float* rgb;
    uint8_t* clamped;
    float* depths;
    uint32_t* tiles_touched;
    uint32_t* cells_touched;
    TileRect* rect;
    uint32_t* sort_key;
    uint32_t* sort_val;
    int32_t* counters;
};

__device__ __forceinline__ float ndc_to_pix(float v, int S) {
    // evaluated in double like the reference (auxiliary.h:41-44)
    return ((v + 1.0) * S - 1.0) * 0.5;
}

__device__ __forceinline__ V3 sh_to_rgb(int deg, const float* __restrict__ sh, float3 p, const float* __restrict__ campos,
                                        unsigned* clamp_bits) {
    V3 pos = {p.x, p.y, p.z};
    V3 cam = {campos[0], campos[1], campos[2]};
    V3 dir = pos - cam;
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len;
    dir.y = dir.y / len;
    dir.z = dir.z / len;

    V3 result = GSR_SH_C0 * ldv3(sh);
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        result = result - GSR_SH_C1 * y * ldv3(sh + 3) + GSR_SH_C1 * z * ldv3(sh + 6) - GSR_SH_C1 * x * ldv3(sh + 9);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
            result = result +
                     GSR_SH_C2_0 * xy * ldv3(sh + 12) +
                     GSR_SH_C2_1 * yz * ldv3(sh + 15) +
                     GSR_SH_C2_2 * (2.0f * zz - xx - yy) * ldv3(sh + 18) +
                     GSR_SH_C2_3 * xz * ldv3(sh + 21) +
                     GSR_SH_C2_4 * (xx - yy) * ldv3(sh + 24);
            if (deg > 2) {
                result = result +
                         GSR_SH_C3_0 * y * (3.0f * xx - yy) * ldv3(sh + 27) +
                         GSR_SH_C3_1 * xy * z * ldv3(sh + 30) +
                         GSR_SH_C3_2 * y * (4.0f * zz - xx - yy) * ldv3(sh + 33) +
                         GSR_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * ldv3(sh + 36) +
                         GSR_SH_C3_4 * x * (4.0f * zz - xx - yy) * ldv3(sh + 39) +
                         GSR_SH_C3_5 * z * (xx - yy) * ldv3(sh + 42) +
                         GSR_SH_C3_6 * x * (xx - 3.0f * yy) * ldv3(sh + 45);
            }
        }
    }
    result.x += 0.5f;
    result.y += 0.5f;
    result.z += 0.5f;
    unsigned bits = 0;
    if (result.x < 0) bits |= 1u;
    if (result.y < 0) bits |= 2u;
    if (result.z < 0) bits |= 4u;
    *clamp_bits = bits;
    result.x = fmaxf(result.x, 0.0f);
    result.y = fmaxf(result.y, 0.0f);
    result.z = fmaxf(result.z, 0.0f);
    return result;
}

__global__ void __launch_bounds__(256) preprocess_fwd_kernel(const __grid_constant__ PreFwdParams p) {
    // camera matrices: one coalesced read per block into shared memory
    __shared__ float s_view[16];
    __shared__ float s_proj[16];
    if (threadIdx.x < 16) s_view[threadIdx.x] = p.view[threadIdx.x];
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = p.proj[threadIdx.x - 16];
    __syncthreads();

    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    // defaults for Gaussians that are not rendered
    int out_radius = 0;
    uint32_t out_tiles = 0, out_cells = 0;
    uint32_t out_key = 0xFFFFFFFFu;   // culled Gaussians sort to the back
    TileRect out_rect = {0, 0, 0, 0};
    bool visible = false;

    const float3 p_orig = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
    const float4 p_hom = xform_point_4x4(p_orig, s_proj);
    const float p_w = 1.0f / (p_hom.w + 0.0000001f);
    const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
    const float3 p_view = xform_point_4x3(p_orig, s_view);

    do {
        if (p_view.z <= NEAR_Z) {
            if (p.prefiltered) atomicExch(&p.counters[0], 1);
            break;
        }

        float cov3D_local[6];
        const float* cov3D;
        if (p.cov3D_precomp != nullptr) {
            cov3D = p.cov3D_precomp + 6 * (size_t)idx;
        } else {
            const float3 sc = {p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]};
            const float4 rot = reinterpret_cast<const float4*>(p.rotations)[idx];
            cov3d_from_scale_rot(sc, p.scale_modifier, rot, cov3D_local);
            cov3D = cov3D_local;
        }

        const Ewa e = ewa_project(p_orig, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, s_view);
        float c00 = e.cov.m[0][0], c01 = e.cov.m[0][1], c11 = e.cov.m[1][1];

        // opacity compensation of the low-pass filter; mixed precision as in forward.cu:112-118
        const float det_0 = max(1e-6, (double)(c00 * c11 - c01 * c01));
        const float det_1 = max(1e-6, (double)((c00 + p.kernel_size) * (c11 + p.kernel_size) - c01 * c01));
        float coef = sqrt(det_0 / (det_1 + 1e-6) + 1e-6);
        if (det_0 <= 1e-6 || det_1 <= 1e-6) coef = 0.0f;
        c00 += p.kernel_size;
        c11 += p.kernel_size;
        const float4 cov = {float(c00), float(c01), float(c11), float(coef)};

        const float det = (cov.x * cov.z - cov.y * cov.y);
        if (det == 0.0f) break;
        const float det_inv = 1.f / det;
        const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

        const float mid = 0.5f * (cov.x + cov.z);
        const float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
        const float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
        const float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
        const float2 point_image = {ndc_to_pix(p_proj.x, p.W), ndc_to_pix(p_proj.y, p.H)};

        // tile rectangle (auxiliary.h:46-56); max_radius is the radius converted to int
        const int max_radius = (int)my_radius;
        const unsigned gx = (unsigned)p.grid_x, gy = (unsigned)p.grid_y;
        unsigned rx0 = min(gx, (unsigned)max((int)0, (int)((point_image.x - max_radius) / TILE)));
        unsigned ry0 = min(gy, (unsigned)max((int)0, (int)((point_image.y - max_radius) / TILE)));
        unsigned rx1 = min(gx, (unsigned)max((int)0, (int)((point_image.x + max_radius + TILE - 1) / TILE)));
        unsigned ry1 = min(gy, (unsigned)max((int)0, (int)((point_image.y + max_radius + TILE - 1) / TILE)));
        if ((rx1 - rx0) * (ry1 - ry0) == 0) break;

        // colour from SH (the colors_precomp path reads colours only while compositing)
        if (p.colors_precomp == nullptr) {
            unsigned bits;
            const V3 c = sh_to_rgb(p.D, p.shs + (size_t)idx * p.M * 3, p_orig, p.campos, &bits);
            p.rgb[3 * idx + 0] = c.x;
            p.rgb[3 * idx + 1] = c.y;
            p.rgb[3 * idx + 2] = c.z;
            p.clamped[idx] = (uint8_t)bits;
        }

        visible = true;
        out_radius = (int)my_radius;
        // shard clip: radii stay those of the full image, only the binned rows shrink
        const unsigned cy0 = max(ry0, (unsigned)p.ty0), cy1 = min(ry1, (unsigned)p.ty1);
        const unsigned rows = cy1 > cy0 ? cy1 - cy0 : 0u;
        out_tiles = rows * (rx1 - rx0);
        out_rect = {(uint16_t)rx0, (uint16_t)(rows ? cy0 : 0u), (uint16_t)rx1, (uint16_t)(rows ? cy1 : 0u)};
        if (rows) out_cells = ((rx1 - 1) / CELL - rx0 / CELL + 1) * ((cy1 - 1) / CELL - cy0 / CELL + 1);
        out_key = __float_as_uint(p_view.z);

        // Screen-space half extents of the region where this Gaussian can reach alpha >= 1/255
        // (axis-aligned box of the ellipse d^T Sigma^-1 d <= 2 ln(255 o)), inflated by a safety margin
        // far above fp32 rounding.  The compositing kernels use it to skip, per warp, Gaussians that
        // cannot touch any of the warp's pixels; the skipped pairs are exactly pairs the reference
        // rejects with its alpha < 1/255 test (forward.cu:365), so results are unchanged.
        // Negative extents: opacity too low to ever pass the test.  NaNs fall through as "keep".
        const float op = p.opacities[idx] * cov.w;
        float hx = -1.f, hy = -1.f;
        if (!(op < (1.0f / 255.0f) * 0.9999f)) {
            const float tau = logf(255.0f * op) * 1.0001f + 1e-4f;
            hx = sqrtf(2.f * tau * cov.x) * 1.0001f + 0.05f;
            hy = sqrtf(2.f * tau * cov.z) * 1.0001f + 0.05f;
        }
        p.rec[2 * (size_t)idx + 0] = make_float4(point_image.x, point_image.y, hx, hy);
        p.rec[2 * (size_t)idx + 1] = make_float4(conic.x, conic.y, conic.z, op);
        p.depths[idx] = p_view.z;
    } while (false);

    p.radii[idx] = out_radius;
    p.tiles_touched[idx] = out_tiles;
    p.cells_touched[idx] = out_cells;
    p.rect[idx] = out_rect;
    p.sort_key[idx] = out_key;
    p.sort_val[idx] = (uint32_t)idx;

    // visible count and total instance count R (one atomic each per warp)
    const unsigned active = __activemask();
    const unsigned ballot = __ballot_sync(active, visible);
    const unsigned warp_tiles = __reduce_add_sync(active, out_tiles);
    if (ballot != 0 && (threadIdx.x & 31) == (__ffs(ballot) - 1)) {
        atomicAdd(&p.counters[1], __popc(ballot));
        atomicAdd(reinterpret_cast<unsigned long long*>(&p.counters[2]), (unsigned long long)warp_tiles);
    }
}

int launch_preprocess_fwd(const GsrForwardArgs& a, const GeomState& g, int ty0, int ty1, cudaStream_t s) {
    PreFwdParams p;
    p.P = a.P; p.D = a.D; p.M = a.M; p.W = a.W; p.H = a.H;
    p.grid_x = tiles_x(a.W); p.grid_y = tiles_y(a.H);
    p.ty0 = ty0; p.ty1 = ty1;
    p.focal_y = a.H / (2.0f * a.tan_fovy);
    p.focal_x = a.W / (2.0f * a.tan_fovx);
    p.tan_fovx = a.tan_fovx; p.tan_fovy = a.tan_fovy;
    p.kernel_size = a.kernel_size; p.scale_modifier = a.scale_modifier;
    p.prefiltered = a.prefiltered;
    p.means3D = a.means3D; p.shs = a.shs; p.colors_precomp = a.colors_precomp;
    p.opacities = a.opacities; p.scales = a.scales; p.rotations = a.rotations;
    p.cov3D_precomp = a.cov3D_precomp; p.campos = a.campos;
    p.view = a.viewmatrix; p.proj = a.projmatrix;
    p.radii = a.radii; p.rec = g.rec; p.rgb = g.rgb; p.clamped = g.clamped; p.depths = g.depths;
    p.tiles_touched = g.tiles_touched; p.cells_touched = g.cells_touched; p.rect = g.rect;
    p.sort_key = g.key_a; p.sort_val = g.val_a; p.counters = g.counters;
    const int blocks = (a.P + 255) / 256;
    preprocess_fwd_kernel<<<blocks, 256, 0, s>>>(p);
    count_launches(1);
    return 0;
}

}  // namespace gsr
