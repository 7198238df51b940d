/*
 * gsrast.h -- C ABI of libgsrast.so, the sm_100a Gaussian-splat rasterizer.
 *
 * This header is the drop-in boundary for the hot path of jkulhanek/wild-gaussians:
 * it replaces the C++ entry points the reference's pybind11 module binds
 *   CudaRasterizer::Rasterizer::forward      (DGR/cuda_rasterizer/rasterizer.h:31-57,  rasterizer_impl.cu:198-340)
 *   CudaRasterizer::Rasterizer::backward     (DGR/cuda_rasterizer/rasterizer.h:59-88,  rasterizer_impl.cu:344-444)
 *   CudaRasterizer::Rasterizer::markVisible  (DGR/cuda_rasterizer/rasterizer.h:24-29,  rasterizer_impl.cu:141-153)
 * (DGR = submodules/diff-gaussian-rasterization) and is what
 *   RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible  (DGR/rasterize_points.cu:35-225)
 * would call instead.  Plain C: raw device pointers, ints and floats, an opaque stream
 * handle (a cudaStream_t passed as void*), no torch types, no exceptions, no allocation
 * inside the library.  Every function returns 0 on success, a positive cudaError_t or a
 * negative GSR_E_* code on failure; gsr_last_error() then holds a thread-local message.
 *
 * All tensors are contiguous fp32 (ints where stated) on the current CUDA device.
 * "absent" optional inputs are NULL (the reference's zero-element tensors,
 * DGR/diff_gaussian_rasterization/__init__.py:218-228).
 *
 * Buffer protocol (replaces the reference's three std::function<char*(size_t)> resize
 * callbacks, rasterize_points.cu:27-33,78-80):
 *   1. gsr_forward_sizes()     -> bytes for geom_buffer and img_buffer      (caller allocates)
 *   2. gsr_forward_geometry()  -> per-Gaussian projection, depth ordering, instance counts
 *                                  (one stream sync, like rasterizer_impl.cu:283-284)
 *   3. gsr_binning_sizes()     -> bytes for binning_buffer and the transient scratch
 *   4. gsr_forward_render()    -> tile binning + per-tile front-to-back composite
 * gsr_forward() does 1-4 in one call through a C allocation callback.
 * geom/binning/img buffers must be kept (unmodified) for gsr_backward(); scratch may be
 * released as soon as gsr_forward_render() has been enqueued (stream-ordered allocators) or
 * completed.
 */
#ifndef GSRAST_H_INCLUDED
#define GSRAST_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 2

/* negative status codes (positive values are cudaError_t) */
#define GSR_E_INVALID      (-1)  /* bad argument (NULL where required, P<0, ...)            */
#define GSR_E_PREFILTERED  (-2)  /* a Gaussian was culled although prefiltered=1
                                    (reference: printf + __trap, auxiliary.h:156-160)        */
#define GSR_E_CHANNELS     (-3)  /* reserved: non-RGB without colors_precomp
                                    (rasterizer_impl.cu:244-247); this build is RGB only     */
#define GSR_E_OVERFLOW     (-4)  /* instance count does not fit the supplied buffers         */

/* which buffer an allocation callback is asked for */
#define GSR_BUF_GEOM     0
#define GSR_BUF_BINNING  1
#define GSR_BUF_IMG      2
#define GSR_BUF_SCRATCH  3

/* Allocation callback for gsr_forward(): must return a device pointer to at least `bytes`
 * bytes, 256-byte aligned, valid at least until gsr_backward() for GEOM/BINNING/IMG.     */
typedef void* (*gsr_alloc_fn)(void* ctx, int which, size_t bytes);

/* Inputs of one forward rasterization.  Field meaning follows
 * CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-225).                        */
typedef struct GsrForwardArgs {
    int P;                        /* number of Gaussians                                    */
    int D;                        /* active SH degree (0..3); ignored with colors_precomp   */
    int M;                        /* SH coefficients per Gaussian, (deg_max+1)^2; 0 if none */
    int W, H;                     /* image size in pixels                                   */
    const float* background;      /* [3]                                                    */
    const float* means3D;         /* [P,3]                                                  */
    const float* shs;             /* [P,M,3] or NULL                                        */
    const float* colors_precomp;  /* [P,3]   or NULL (exactly one of shs/colors_precomp)    */
    const float* opacities;       /* [P] (or [P,1])                                         */
    const float* scales;          /* [P,3]   or NULL                                        */
    float        scale_modifier;
    const float* rotations;       /* [P,4] (r,x,y,z), used un-normalised; or NULL           */
    const float* cov3D_precomp;   /* [P,6]   or NULL (exactly one of scales+rot / cov3D)    */
    const float* viewmatrix;      /* [16] row-vector convention (translation at 12..14)     */
    const float* projmatrix;      /* [16] view*proj, same convention                        */
    const float* campos;          /* [3]                                                    */
    float tan_fovx, tan_fovy;
    float kernel_size;            /* Mip-Splatting 2D low-pass (forward.cu:108-121)         */
    const float* subpixel_offset; /* [H,W,2]                                                */
    int prefiltered;
    int debug;                    /* !=0: synchronise + check after every stage             */
    /* Screen-space shard (multi-GPU tile-row partition): only tile rows
     * [tile_y0, tile_y1) are binned and composited; pixels of other rows are not written.
     * tile_y0 = tile_y1 = 0 means the whole image.                                         */
    int tile_y0, tile_y1;
    /* outputs */
    float* out_color;             /* [3,H,W]  (rows of the shard are written)               */
    int*   radii;                 /* [P]      screen radius in px, 0 = not rendered         */
    /* Multi-GPU image all-gather fused into the composite (optional): `peer_images` is a DEVICE array of
     * `n_peer_images` pointers, one per rank (this rank included), each to a [4,H,W] fp32 image (colour planes +
     * final transmittance) in peer-mapped memory (NVLink / NVSwitch symmetric memory).  The composite then stores
     * every pixel of its band into ALL of them instead of out_color (which may be NULL); after a barrier across the
     * ranks every rank holds the full image.  NULL / 0: single-GPU behaviour.                                   */
    const void* const* peer_images;
    int    n_peer_images;
} GsrForwardArgs;

/* Inputs/outputs of the backward pass.  Field meaning follows
 * CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:344-377).
 * Every output array is written completely by the call (zeros for Gaussians that were not
 * rendered): no pre-zeroing is required, unlike rasterize_points.cu:157-165.               */
typedef struct GsrBackwardArgs {
    int P, D, M;
    int R;                        /* num_rendered returned by the forward                   */
    int W, H;
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    float        scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy;
    float kernel_size;
    const float* subpixel_offset;
    const int*   radii;           /* [P] from the forward                                   */
    const void*  geom_buffer;
    const void*  binning_buffer;
    const void*  img_buffer;
    const float* dL_dpix;         /* [3,H,W] upstream gradient of out_color                 */
    int debug;
    int tile_y0, tile_y1;         /* same shard as the forward                              */
    void*  accum_scratch;         /* gsr_backward_scratch_bytes(P) bytes.  gsr_backward / _partials zero it first unless
                                     accum_is_zero != 0; gsr_backward_finalize leaves it ZEROED again (a stream-
                                     ordered fill after its kernel), so a caller that keeps the buffer skips the fill. */
    int    accum_is_zero;
    /* outputs */
    float* dL_dmean2D;            /* [P,3]  x,y: NDC-scaled screen grad; z: sum |gx|+|gy|
                                     (backward.cu:590-595)                                  */
    float* dL_dconic;             /* [P,4]  (x,y,-,w) as backward.cu:598-600; may be NULL   */
    float* dL_dopacity;           /* [P]                                                    */
    float* dL_dcolor;             /* [P,3]                                                  */
    float* dL_dmean3D;            /* [P,3]                                                  */
    float* dL_dcov3D;             /* [P,6]; may be NULL when cov3D_precomp is NULL          */
    float* dL_dsh;                /* [P,M,3]; required iff shs != NULL                      */
    float* dL_dscale;             /* [P,3];  required iff scales != NULL                    */
    float* dL_drot;               /* [P,4];  required iff rotations != NULL                 */
} GsrBackwardArgs;

/* Counters of the last forward on this geom buffer (device->host copied on request).     */
typedef struct GsrStats {
    int      num_rendered;        /* R: (Gaussian, tile) instances                          */
    int      num_visible;         /* V: Gaussians with radii > 0                            */
    int      num_tiles;           /* tiles in the shard                                     */
    int      num_coarse;          /* N1: (Gaussian, 8x8-tile cell) items of the two-level binning */
} GsrStats;

int         gsr_abi_version(void);
const char* gsr_last_error(void);

/* -- forward ----------------------------------------------------------------------------- */
int gsr_forward_sizes(int P, int M, int W, int H, size_t* geom_bytes, size_t* img_bytes);
/* num_rendered = R, the number of (Gaussian, tile) instances (what the reference returns);
 * num_coarse   = number of (Gaussian, 8x8-tile cell) items of the two-level binning: it sizes the
 *                transient scratch and must be passed back to gsr_binning_sizes / gsr_forward_render. */
int gsr_forward_geometry(const GsrForwardArgs* args, void* geom_buffer, void* img_buffer,
                         void* stream, int* num_rendered, int* num_coarse);
int gsr_binning_sizes(int P, int W, int H, int num_rendered, int num_coarse,
                      size_t* binning_bytes, size_t* scratch_bytes);
int gsr_forward_render(const GsrForwardArgs* args, void* geom_buffer, void* img_buffer,
                       void* binning_buffer, void* scratch, int num_rendered, int num_coarse,
                       void* stream);
int gsr_forward(const GsrForwardArgs* args, gsr_alloc_fn alloc, void* alloc_ctx,
                void* stream, int* num_rendered);
/* The whole forward WITHOUT the host synchronisation in its middle (the reference blocks on the instance count at
 * rasterizer_impl.cu:283-284, gsr_forward_geometry does the same): the caller supplies the binning buffer and the
 * scratch sized for `rendered_capacity` instances / `coarse_capacity` coarse items (gsr_binning_sizes; typically the
 * previous frame's counts plus a margin), every launch configuration is derived from the capacities and the actual
 * counts stay on the device.  If the scene does not fit, the binning / composite stages turn into no-ops (empty tile
 * lists, no out-of-bounds access) and an overflow flag is raised.  gsr_forward_status() -- one stream synchronisation,
 * whenever the caller needs the counts -- returns the exact num_rendered / num_coarse (valid even after an overflow, so
 * one retry with exactly sized buffers, e.g. through gsr_forward_geometry + gsr_forward_render, always succeeds) and
 * overflow != 0 if the forward must be repeated.  CUDA-graph capturable (no synchronisation, fixed launch shapes).    */
int gsr_forward_async(const GsrForwardArgs* args, void* geom_buffer, void* img_buffer, void* binning_buffer,
                      int rendered_capacity, void* scratch, int coarse_capacity, void* stream);
int gsr_forward_status(const void* geom_buffer, int P, int M, void* stream, int* num_rendered, int* num_coarse,
                       int* overflow);

/* Re-composite with different per-Gaussian colours on the SAME geometry/binning state
 * (wild-gaussians renders raw + appearance-toned colours per step, method.py:1573-1611).
 * Writes out_color (and final_T / n_contrib into img_buffer2, sized like img_buffer).     */
int gsr_forward_recolor(const GsrForwardArgs* args, const void* geom_buffer,
                        const void* binning_buffer, const void* img_buffer,
                        void* img_buffer2, void* stream);

/* -- backward ---------------------------------------------------------------------------- */
size_t gsr_backward_scratch_bytes(int P);
int    gsr_backward(const GsrBackwardArgs* args, void* stream);
/* The two halves of gsr_backward(), for the tile-row sharded (multi-GPU) path where the
 * per-Gaussian partial sums of all shards are added (NCCL all-reduce over accum_scratch, a dense
 * [P][12] fp32 array starting at the first 256-byte aligned address) between them:
 *   partials: zero accum_scratch, run the per-tile backward composite of this shard into it
 *             (needs dL_dpix, img/binning buffers; writes no output tensor)
 *   finalize: per-Gaussian chain rule from accum_scratch to every output tensor                     */
int    gsr_backward_partials(const GsrBackwardArgs* args, void* stream);
int    gsr_backward_finalize(const GsrBackwardArgs* args, void* stream);
/* Reduction-fused variant of gsr_backward_partials for ranks that share a symmetric accumulator (one [P][12] fp32
 * array per rank, each mapped into every process; NVLink/NVSwitch peer memory): the per-tile backward composite adds
 * its sums directly into the accumulators of ALL ranks -- through `multicast_accum` (NVSwitch multicast address: one
 * multimem.red per 16 bytes, the switch updates every replica) when it is non-NULL, else through the `n_peers` device
 * pointers stored in the DEVICE array `peer_accum_dev`.  The arrays must have been zeroed on every rank (and a
 * barrier passed) before any rank calls this; after a second barrier every rank holds the complete sums and calls
 * gsr_backward_finalize with accum_scratch = its own array.  Replaces partials + all-reduce.                       */
int    gsr_backward_partials_peers(const GsrBackwardArgs* args, const void* const* peer_accum_dev, int n_peers,
                                   void* multicast_accum, void* stream);

/* "Pull" form of the multi-GPU reduction (GSR_PEER_REDUCE=3): no remote atomics at all.  gsr_backward_partials_marked adds
 * this rank's band into ITS OWN accumulator (args->accum_scratch) and sets touched[i] = 1 for every Gaussian i it added to
 * (one byte per Gaussian, all-zero on entry).  After a barrier, gsr_backward_finalize_pull forms each Gaussian's complete
 * sums as own row + the rows the other ranks marked, read through their peer-mapped accumulators in rank order 0..n-1
 * (the same order on every rank: the resulting gradients are bit-identical across ranks), and runs the chain rule.
 * peer_accums / peer_touched: HOST arrays of n_peers (<= GSR_MAX_PULL_PEERS) device pointers -- the ranks' buffers as mapped
 * into this process (entry `self` = this rank's own buffers).
 * The buffers of THIS pass must stay untouched until every rank has finished reading them, i.e. until the barrier of the
 * next pass: keep two sets and pass the previous pass's set as clear_accum / clear_touched (may be NULL) -- its marked
 * rows and marks are zeroed by this call.                                                                             */
int    gsr_backward_partials_marked(const GsrBackwardArgs* args, unsigned char* touched, void* stream);
#define GSR_MAX_PULL_PEERS 8
int    gsr_backward_finalize_pull(const GsrBackwardArgs* args, const void* const* peer_accums,
                                  const void* const* peer_touched, int n_peers, int self, void* clear_accum,
                                  unsigned char* clear_touched, void* stream);

/* -- markVisible (rasterizer_impl.cu:54-66,141-153): present[i] = (view*p).z > 0.2 ------- */
int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* -- introspection used by parity tests and the benchmark -------------------------------- */
/* accum_alpha (= final transmittance T per pixel, [H,W] f32) lives at the first
 * 128-byte-aligned address of img_buffer, as in the reference ImageState
 * (rasterizer_impl.cu:172-178; read by __init__.py:101-113).                               */
int gsr_img_views(const void* img_buffer, int W, int H,
                  const float** final_T, const uint32_t** n_contrib, const uint32_t** ranges /* uint2[T] */);
/* sorted (tile-major, depth-minor) Gaussian index list, R entries                          */
int gsr_binning_views(const void* binning_buffer, int num_rendered, const uint32_t** point_list);
/* per-Gaussian projected state: depths f32[P], packed records float4[2P]
 * {x, y, hx, hy | conic.x, conic.y, conic.z, opacity*coef} (hx, hy: conservative half extents of the
 * alpha >= 1/255 footprint, used for per-warp culling), tiles_touched u32[P], rgb f32[3P]            */
int gsr_geom_views(const void* geom_buffer, int P, int M, const float** depths,
                   const float** records, const uint32_t** tiles_touched, const float** rgb);
int gsr_get_stats(const void* geom_buffer, int P, int M, void* stream, GsrStats* out);

/* -- fused per-Gaussian colour op of wild-gaussians (optional entry points) ---------------------------
 * Replaces, for sh_degree = 3 (48 SH features), 24 per-Gaussian embedding features and a 32-d image embedding, the
 * PyTorch statements of wildgaussians/method.py that turn the model's features into the two colour sets the rasterizer
 * composites: `features.clamp_max(1)` (:1570), eval_sh + 0.5 + clamp_min(0) for the raw colours (:1571-1579, :493-548),
 * EmbeddingModel.forward (59 -> 128 -> 128 -> 6 MLP, x 0.01, affine on the features; :874-900) and the toned colours
 * (:1586-1598).  The MLP runs on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
 *   1. gsr_appearance_pack_weights(): fp32 nn.Linear parameters + the image embedding -> packed bf16 operand image
 *      (gsr_appearance_packed_weight_bytes() bytes, 16-byte aligned); once per step
 *   2. gsr_appearance_colors_forward():  colors_raw (optional) and colors_toned, [P,3] each
 *   3. gsr_appearance_colors_backward(): gradients of every per-Gaussian input; weight gradients are accumulated in
 *      `grad_pack` (gsr_appearance_grad_pack_bytes() bytes, zeroed by the call)
 *   4. gsr_appearance_unpack_grads():    grad_pack -> dW1 [128,59], db1 [128], dW2 [128,128], db2 [128], dW3 [6,128],
 *      db3 [6], d(image embedding) [32]
 * `status` (optional, device int, zero it first) becomes non-zero if an internal barrier wait timed out.              */
typedef struct GsrAppearanceArgs {
    int P;
    int sh_degree;                    /* ACTIVE degree 0..3 (method.py: active_sh_degree); storage is always degree 3 */
    const float* features_dc;         /* [P,3]   raw (unclamped) DC features                                          */
    const float* features_rest;       /* [P,45]  raw higher-order features, coefficient-major (k, channel)            */
    const float* embeddings;          /* [P,24]  per-Gaussian appearance features                                      */
    const float* means3D;             /* [P,3]                                                                         */
    const float* campos;              /* [3]                                                                           */
    const void*  packed_weights;      /* from gsr_appearance_pack_weights                                              */
    float* colors_raw;                /* [P,3] out, may be NULL                                                        */
    float* colors_toned;              /* [P,3] out                                                                     */
    /* backward only */
    const float* dL_dcolors_raw;      /* [P,3] or NULL                                                                 */
    const float* dL_dcolors_toned;    /* [P,3]                                                                         */
    float* dL_dfeatures_dc;           /* [P,3]                                                                         */
    float* dL_dfeatures_rest;         /* [P,45]                                                                        */
    float* dL_dembeddings;            /* [P,24]                                                                        */
    float* dL_dmeans3D;               /* [P,3]  through the view direction of the SH evaluation                        */
    float* grad_pack;                 /* packed weight-gradient image                                                  */
    int*   status;                    /* optional                                                                      */
} GsrAppearanceArgs;

size_t gsr_appearance_packed_weight_bytes(void);
size_t gsr_appearance_grad_pack_bytes(void);
int gsr_appearance_pack_weights(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                                const float* b3, const float* appearance_embedding, void* packed_weights, void* stream);
int gsr_appearance_colors_forward(const GsrAppearanceArgs* args, void* stream);
int gsr_appearance_colors_backward(const GsrAppearanceArgs* args, void* stream);
int gsr_appearance_unpack_grads(const float* grad_pack, const float* W1, const float* appearance_embedding, float* dW1,
                                float* db1, float* dW2, float* db2, float* dW3, float* db3,
                                float* dappearance_embedding, void* stream);

/* -- parameter activations + 3D filter (optional entry points) --------------------------------------------
 * GaussianModel.get_gaussians (wildgaussians/method.py:1060-1086) in one kernel per direction: rotations = normalize(raw),
 * scales = sqrt(exp(raw)^2 + filter_3D^2), opacities = sigmoid(raw) * sqrt(prod(exp(raw)^2) / prod(exp(raw)^2 + filter_3D^2)).
 * All arrays fp32, [P,3] / [P] / [P,4] (16-byte aligned) / [P].  The backward recomputes the forward; upstream gradients may be
 * NULL (treated as zero).                                                                                                  */
int gsr_gaussian_activations_forward(int P, const float* scales_raw, const float* opacities_raw, const float* rotations_raw,
                                     const float* filter_3D, float* scales, float* opacities, float* rotations, void* stream);
int gsr_gaussian_activations_backward(int P, const float* scales_raw, const float* opacities_raw, const float* rotations_raw,
                                      const float* filter_3D, const float* dL_dscales, const float* dL_dopacities,
                                      const float* dL_drotations, float* dL_dscales_raw, float* dL_dopacities_raw,
                                      float* dL_drotations_raw, void* stream);

/* -- densification statistics (optional entry point) -----------------------------------------------------
 * One pass over the Gaussians replacing wildgaussians/method.py:1997-1998 + GaussianModel.add_densification_stats
 * (:1470-1477): for every Gaussian with radii > 0: max_radii2D = max(., radii); xyz_grad += |grad.xy|;
 * accum_abs += |grad.z|; accum_abs_max = max(., |grad.z|) (both NULL when use_gof_abs_gradient is off); denom += 1.
 * viewspace_grad is the [P,3] gradient the rasterizer leaves in means2D.grad (summed over the step's passes).          */
int gsr_densification_stats(int P, const int* radii, const float* viewspace_grad, float* max_radii2D, float* xyz_grad,
                            float* xyz_gradient_accum_abs, float* xyz_gradient_accum_abs_max, float* denom, void* stream);

/* -- optimizer step (optional entry point) ------------------------------------------------------------------
 * One launch replacing `self.model.optimizer.step()` (wildgaussians/method.py:2019) of the Adam optimizer built at
 * method.py:1033-1049 (lr per group, eps = 1e-15, weight_decay on the appearance-embedding table only; amsgrad /
 * maximize off).  The algorithm lives in the reference's third-party dependency PyTorch (optim/adam.py,
 * _multi_tensor_adam); this entry point performs the same fp32 operations per element, all tensors in one pass.
 * `step` is the 1-based step count of the tensor AFTER this update (PyTorch increments state["step"] first); lr / step
 * may differ per segment.  All pointers are device pointers to fp32 arrays of `n` elements; `grad` is only written when
 * zero_grads != 0 (it is then left all-zero, the in-place counterpart of zero_grad()).                                  */
#define GSR_ADAM_MAX_SEGMENTS 32
typedef struct GsrAdamSegment {
    float* param;
    float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long n;
    long long step;
    double lr;
    double weight_decay;
} GsrAdamSegment;
int gsr_adam_step(const GsrAdamSegment* segments, int num_segments, double beta1, double beta2, double eps, int zero_grads,
                  void* stream);

/* -- 3D filter sizes over all training cameras (optional entry point) -----------------------------------------
 * Replaces GaussianModel.compute_3D_filter (wildgaussians/method.py:1140-1190): per Gaussian the smallest camera-space depth
 * over the cameras that see it (depth > 0.2 and projection inside the image enlarged by 15 %), Gaussians no camera sees get
 * the largest such depth; filter_3D = distance / focal_length * sqrt(0.2).  `cameras` is a device array of num_cameras
 * records of 20 floats: R[9] (row-major, the matrix method.py:1165 multiplies with: xyz @ R + T), T[3], fx, fy, W/2, H/2,
 * -0.15 W, 1.15 W, -0.15 H, 1.15 H; `focal_length` = the largest fx (method.py:1178-1179).  `scratch`: device memory of
 * gsr_filter3d_scratch_bytes(P) bytes, 16-byte aligned.  filter_3D: [P] (the reference stores it as [P,1]).               */
size_t gsr_filter3d_scratch_bytes(int P);
int gsr_compute_3d_filter(int P, const float* xyz, int num_cameras, const float* cameras, float focal_length,
                          float* filter_3D, void* scratch, void* stream);

/* -- initial scales from the point cloud (optional entry point) -------------------------------------------
 * Replaces distCUDA2 of the reference's simple-knn submodule (spatial.cu:15-26 -> SimpleKNN::knn, simple_knn.cu:185-220;
 * called once at method.py:1001): mean_dist2[i] = mean of the squared distances from point i to its 3 nearest OTHER points
 * (fp32, the reference's expression and summation order; exact search, so the values are the reference's).
 * points: [P,3] fp32 on the device; scratch: gsr_knn_scratch_bytes(P) bytes of device memory, 256-byte aligned.           */
size_t gsr_knn_scratch_bytes(int P);
int gsr_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* scratch, void* stream);

/* Optional per-stage device timing (cudaEvents on the caller's stream around each stage of the
 * next forward / backward calls).  The caller synchronises the stream, then reads the stage times of
 * the most recent calls in milliseconds (-1 for stages that did not run).  Not thread-safe.        */
void        gsr_profile_enable(int on);
int         gsr_profile_stage_count(void);
const char* gsr_profile_stage_name(int i);
int         gsr_profile_read(float* ms, int n);
/* number of kernels this library has launched in this process                                     */
unsigned long long gsr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GSRAST_H_INCLUDED */
