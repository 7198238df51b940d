// adam.cu -- the optimizer step of wild-gaussians' train iteration in ONE launch.
//
// What it replaces: `self.model.optimizer.step()` (wildgaussians/method.py:2019) on the optimizer built at
// method.py:1033-1049 -- torch.optim.Adam(groups, lr=1.0, eps=1e-15) over the per-Gaussian tensors (xyz,
// features_dc, opacities, scales, rotations, embeddings, features_rest), the appearance-embedding table
// (weight_decay = appearance_embedding_regularization) and the six tensors of the appearance MLP.  torch runs its
// "foreach" path: nine multi-tensor launches (lerp, mul, addcmul, sqrt, div, add, addcdiv, ...), each streaming the
// 83 floats per Gaussian again, i.e. about 3x the compulsory traffic.  Here every parameter element is read once
// (param, grad, exp_avg, exp_avg_sq) and written once (param, exp_avg, exp_avg_sq): 28 B per element, HBM-bound.
//
// Arithmetic: exactly the sequence of fp32 operations of torch/optim/adam.py:_multi_tensor_adam (pinned: the torch
// 2.x line the reference depends on; amsgrad / maximize / capturable off, as the reference leaves them):
//     g   = grad + weight_decay * param                      (only when weight_decay != 0)
//     m   = m + (1 - beta1) * (g - m)                        (lerp, weight < 0.5 form, one fma)
//     v   = v * beta2;  v = v + (1 - beta2) * (g * g)        (mul, then foreach addcmul: one fma)
//     d   = sqrt(v) / sqrt(1 - beta2^t) + eps
//     p   = p + (-lr / (1 - beta1^t)) * (m / d)              (addcdiv: one fma)
// The scalar factors (1 - beta1, 1 - beta2, sqrt(1 - beta2^t), -lr / (1 - beta1^t)) are formed on the host in
// double precision and rounded to fp32 once, as torch does when it hands Python floats to its foreach kernels.
//
// Layout: up to GSR_ADAM_MAX_SEGMENTS tensors per launch, described by value in the kernel parameters; the flat
// element space of all segments is cut into chunks of 4096 elements, a CTA owns one chunk (linear scan over at most
// 32 chunk prefixes), 128-bit accesses when the four pointers of the segment are 16-byte aligned.
#include "common.cuh"

namespace gsr {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = 4096;     // elements per CTA: 4 float4 per thread

struct AdamSeg {
    float* param;
    float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long n;
    float neg_step_size;            // -lr / (1 - beta1^t)
    float bias_correction2_sqrt;    // sqrt(1 - beta2^t)
    float weight_decay;
    int vec;                        // 1: all four pointers 16-byte aligned
};

struct AdamParams {
    AdamSeg seg[GSR_ADAM_MAX_SEGMENTS];
    unsigned chunk_end[GSR_ADAM_MAX_SEGMENTS];   // inclusive prefix of the segments' chunk counts
    int num_segments;
    float one_minus_beta1, beta2, one_minus_beta2, eps;
    int zero_grads;
};

__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, const AdamSeg& s, const AdamParams& a) {
    float gg = g;
    if (s.weight_decay != 0.f) gg = fmaf(s.weight_decay, p, gg);       // grad.add(param, alpha = weight_decay)
    m = fmaf(a.one_minus_beta1, gg - m, m);                            // exp_avg.lerp_(grad, 1 - beta1)
    v = __fmul_rn(v, a.beta2);                                         // exp_avg_sq.mul_(beta2)
    v = fmaf(a.one_minus_beta2, __fmul_rn(gg, gg), v);                 //   ._foreach_addcmul_(grad, grad, 1 - beta2): v + value * (g * g)
    const float d = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bias_correction2_sqrt), a.eps);
    p = fmaf(s.neg_step_size, __fdiv_rn(m, d), p);                     // param.addcdiv_(exp_avg, denom, value = -step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_step_kernel(const __grid_constant__ AdamParams a) {
    // which segment owns this CTA's chunk?
    const unsigned c = blockIdx.x;
    int si = 0;
    while (si < a.num_segments - 1 && c >= a.chunk_end[si]) ++si;
    const AdamSeg& s = a.seg[si];
    const unsigned first = si == 0 ? 0u : a.chunk_end[si - 1];
    const long long base = (long long)(c - first) * ADAM_CHUNK;
    const long long end = min(base + (long long)ADAM_CHUNK, s.n);

    if (s.vec && end - base == ADAM_CHUNK) {
        float4* P4 = reinterpret_cast<float4*>(s.param + base);
        float4* G4 = reinterpret_cast<float4*>(s.grad + base);
        float4* M4 = reinterpret_cast<float4*>(s.exp_avg + base);
        float4* V4 = reinterpret_cast<float4*>(s.exp_avg_sq + base);
        float4 p[4], g[4], m[4], v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // all loads first: 16 independent 128-bit requests per thread in flight
            const int i = k * ADAM_THREADS + threadIdx.x;
            p[k] = P4[i]; g[k] = G4[i]; m[k] = M4[i]; v[k] = V4[i];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = k * ADAM_THREADS + threadIdx.x;
            adam_update(p[k].x, g[k].x, m[k].x, v[k].x, s, a);
            adam_update(p[k].y, g[k].y, m[k].y, v[k].y, s, a);
            adam_update(p[k].z, g[k].z, m[k].z, v[k].z, s, a);
            adam_update(p[k].w, g[k].w, m[k].w, v[k].w, s, a);
            P4[i] = p[k]; M4[i] = m[k]; V4[i] = v[k];
            if (a.zero_grads) G4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
        float p = s.param[i], g = s.grad[i], m = s.exp_avg[i], v = s.exp_avg_sq[i];
        adam_update(p, g, m, v, s, a);
        s.param[i] = p; s.exp_avg[i] = m; s.exp_avg_sq[i] = v;
        if (a.zero_grads) s.grad[i] = 0.f;
    }
}

}  // namespace gsr

extern "C" int gsr_adam_step(const GsrAdamSegment* segments, int num_segments, double beta1, double beta2, double eps,
                             int zero_grads, void* stream) {
    using namespace gsr;
    if (num_segments < 0 || num_segments > GSR_ADAM_MAX_SEGMENTS) { set_error("num_segments must be in [0, %d]", GSR_ADAM_MAX_SEGMENTS); return GSR_E_INVALID; }
    if (num_segments == 0) return 0;
    if (!segments) { set_error("segments is NULL"); return GSR_E_INVALID; }
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) { set_error("bad beta1 / beta2 / eps"); return GSR_E_INVALID; }
    AdamParams a{};
    unsigned chunks = 0;
    int k = 0;
    for (int i = 0; i < num_segments; ++i) {
        const GsrAdamSegment& g = segments[i];
        if (g.n < 0 || g.step < 1) { set_error("segment %d: bad n / step", i); return GSR_E_INVALID; }
        if (g.n == 0) continue;
        if (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq) { set_error("segment %d: NULL pointer", i); return GSR_E_INVALID; }
        const long long nchunks = (g.n + ADAM_CHUNK - 1) / ADAM_CHUNK;
        if (nchunks + chunks > 0x7FFFFFFFll) { set_error("too many elements"); return GSR_E_INVALID; }
        AdamSeg& s = a.seg[k];
        s.param = g.param; s.grad = g.grad; s.exp_avg = g.exp_avg; s.exp_avg_sq = g.exp_avg_sq; s.n = g.n;
        // host-side scalars in double, rounded to fp32 once (torch/optim/adam.py: bias_correction1/2, step_size)
        const double bc1 = 1.0 - pow(beta1, (double)g.step), bc2 = 1.0 - pow(beta2, (double)g.step);
        s.neg_step_size = (float)(-(g.lr / bc1));
        s.bias_correction2_sqrt = (float)sqrt(bc2);
        s.weight_decay = (float)g.weight_decay;
        s.vec = ((reinterpret_cast<uintptr_t>(g.param) | reinterpret_cast<uintptr_t>(g.grad) | reinterpret_cast<uintptr_t>(g.exp_avg) |
                  reinterpret_cast<uintptr_t>(g.exp_avg_sq)) & 15u) == 0;
        chunks += (unsigned)nchunks;
        a.chunk_end[k] = chunks;
        ++k;
    }
    if (k == 0) return 0;
    a.num_segments = k;
    a.one_minus_beta1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.one_minus_beta2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.zero_grads = zero_grads;
    adam_step_kernel<<<chunks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(a);
    count_launches(1);
    GSR_CUDA(cudaGetLastError());
    return 0;
}
