// umma_probe.cu -- bring-up probe for the hand-written tcgen05 path of csrc/appearance.cu.
// Not part of the product: a standalone binary that issues single-CTA tcgen05.mma (kind::f16, bf16 operands, fp32
// accumulators in TMEM) on operands laid out the way appearance.cu lays them out in shared memory, for every
// operand-major combination the forward / dgrad / wgrad GEMMs need, and compares with a CPU product.
//   build:  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/umma_probe.cu -o tools/_build/umma_probe
// The shared-memory matrix descriptor fields follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Case {
    uint32_t idesc;          // instruction descriptor (upper 32 bits of the 64-bit idesc operand)
    uint32_t a_lbo, a_sbo;   // bytes
    uint32_t b_lbo, b_sbo;   // bytes
    uint32_t a_kstep, b_kstep;   // bytes added to the start address per K = 16 step
    uint32_t ksteps;
    uint32_t n;              // N of the MMA (columns read back)
    uint32_t a_bytes, b_bytes;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // version_ = 1 (Blackwell)
    // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
    return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* d_out, Case c) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + 65536;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < c.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(a_img + i);
    for (uint32_t i = tid * 16; i < c.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(b_img + i);
    if (tid == 0) {
        const uint32_t b32 = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy writes of the operands must be visible to the async proxy (tcgen05.mma reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sa), b0 = (uint32_t)__cvta_generic_to_shared(sb);
        for (uint32_t k = 0; k < c.ksteps; ++k) {
            const uint64_t da = make_desc(a0 + k * c.a_kstep, c.a_lbo, c.a_sbo);
            const uint64_t db = make_desc(b0 + k * c.b_kstep, c.b_lbo, c.b_sbo);
            const uint32_t acc = k > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(c.idesc), "r"(acc) : "memory");
        }
        const uint32_t b32 = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b32) : "memory");
    }
    {   // wait for the MMAs (phase 0)
        const uint32_t b32 = (uint32_t)__cvta_generic_to_shared(&bar);
        uint32_t done = 0;
        int spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(b32), "r"(0u) : "memory");
            if (++spins > (1 << 22)) { if (tid == 0) printf("probe: mbarrier wait timed out\n"); break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D: row m of the 128 x N accumulator lives in TMEM lane m; warp w may read lanes 32 w .. 32 w + 31
    for (uint32_t col = 0; col < c.n; col += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + col;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) d_out[(size_t)tid * 256 + col + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

static uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                 // c_format = F32
    d |= 1u << 7;                 // a_format = BF16
    d |= 1u << 10;                // b_format = BF16
    d |= (uint32_t)a_mn << 15;    // a_major: 0 = K, 1 = MN
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

static uint16_t bf(float x) { __nv_bfloat16 h = __float2bfloat16(x); uint16_t r; memcpy(&r, &h, 2); return r; }

// Operand storage used by appearance.cu for every activation / weight matrix X[rows][cols] (bf16):
//   byte address(row, col) = (col / 8) * (ROWS * 16) + row * 16 + (col % 8) * 2
// "rows-major-16B": 8 consecutive columns of one row are one 16-byte unit, consecutive rows are consecutive units.
// Read with K = cols it is the canonical K-major INTERLEAVE layout (SBO = 128 B between 8-row groups, LBO = ROWS*16 B
// between the two 8-column halves of a K = 16 step); read with K = rows it is the canonical MN-major INTERLEAVE layout
// (SBO = ROWS*16 B between 8-column chunks of MN, LBO = 128 B between 8-row groups of K).
static void store_mat(std::vector<uint8_t>& img, const std::vector<float>& X, int rows, int cols) {
    img.assign((size_t)rows * cols * 2, 0);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const uint16_t v = bf(X[(size_t)r * cols + c]);
            memcpy(&img[(size_t)(c / 8) * rows * 16 + (size_t)r * 16 + (c % 8) * 2], &v, 2);
        }
}

int main() {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    uint8_t *da, *db; float* dd;
    CK(cudaMalloc(&da, 65536)); CK(cudaMalloc(&db, 65536)); CK(cudaMalloc(&dd, 128 * 256 * 4));
    srand(1);
    auto rnd = [](int rows, int cols) { std::vector<float> v((size_t)rows * cols); for (auto& x : v) x = (float)((rand() % 9) - 4); return v; };
    int fails = 0;
    auto run = [&](const char* name, const std::vector<uint8_t>& ia, const std::vector<uint8_t>& ib, Case c, const std::vector<float>& ref) {
        c.a_bytes = (uint32_t)ia.size(); c.b_bytes = (uint32_t)ib.size();
        CK(cudaMemcpy(da, ia.data(), ia.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, ib.data(), ib.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dd, 0xFF, 128 * 256 * 4));
        probe_kernel<<<1, 128, 131072>>>(da, db, dd, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(3); }
        std::vector<float> out(128 * 256);
        CK(cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int m = 0; m < 128; ++m)
            for (uint32_t n = 0; n < c.n; ++n) {
                const double d = fabs((double)out[(size_t)m * 256 + n] - ref[(size_t)m * c.n + n]);
                if (!(d <= worst)) worst = d;
            }
        printf("%-44s lboA %5u sboA %5u lboB %5u sboB %5u  max|err| = %g %s\n", name, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, worst,
               worst == 0 ? "OK" : "MISMATCH");
        return worst == 0;
    };

    // ---- 1. forward-type GEMM: D[128 x N] = X[128 x K] . W[N x K]^T, both operands K-major -----------------------
    for (int N : {128, 16}) {
        const int K = 64;
        auto X = rnd(128, K), W = rnd(N, K);
        std::vector<float> ref((size_t)128 * N, 0.f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += X[m * K + k] * W[n * K + k]; ref[m * N + n] = s; }
        std::vector<uint8_t> ia, ib; store_mat(ia, X, 128, K); store_mat(ib, W, N, K);
        Case c{}; c.idesc = idesc(128, N, 0, 0); c.ksteps = K / 16; c.n = N;
        c.a_kstep = 2 * 128 * 16; c.b_kstep = 2 * N * 16;
        char nm[64];
        snprintf(nm, sizeof nm, "K-major x K-major N=%d (expected)", N);
        c.a_lbo = 128 * 16; c.a_sbo = 128; c.b_lbo = N * 16; c.b_sbo = 128;
        const bool ok = run(nm, ia, ib, c, ref);
        snprintf(nm, sizeof nm, "K-major x K-major N=%d (lbo/sbo swapped)", N);
        c.a_lbo = 128; c.a_sbo = 128 * 16; c.b_lbo = 128; c.b_sbo = N * 16;
        const bool ok2 = run(nm, ia, ib, c, ref);
        if (!ok) ++fails;
        (void)ok2;
    }
    // ---- 2. wgrad-type GEMM: D[128 x N] = A[rows x 128]^T . B[rows x N], K = rows = 128, both MN-major --------
    for (int N : {128, 64, 16}) {
        const int R = 128;
        auto A = rnd(R, 128), B = rnd(R, N);
        std::vector<float> ref((size_t)128 * N, 0.f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < R; ++k) s += A[k * 128 + m] * B[k * N + n]; ref[m * N + n] = s; }
        std::vector<uint8_t> ia, ib; store_mat(ia, A, R, 128); store_mat(ib, B, R, N);
        Case c{}; c.idesc = idesc(128, N, 1, 1); c.ksteps = R / 16; c.n = N;
        c.a_kstep = 16 * 16; c.b_kstep = 16 * 16;       // 16 rows further
        char nm[64];
        snprintf(nm, sizeof nm, "MN-major x MN-major N=%d (expected)", N);
        c.a_lbo = 128; c.a_sbo = R * 16; c.b_lbo = 128; c.b_sbo = R * 16;
        const bool ok = run(nm, ia, ib, c, ref);
        snprintf(nm, sizeof nm, "MN-major x MN-major N=%d (lbo/sbo swapped)", N);
        c.a_lbo = R * 16; c.a_sbo = 128; c.b_lbo = R * 16; c.b_sbo = 128;
        run(nm, ia, ib, c, ref);
        if (!ok) ++fails;
    }
    // ---- 3. dgrad-type GEMM: D[128 x N] = G[128 x K] . W[K x N], A K-major, B MN-major (W stored [K rows][N cols]) --
    for (int N : {128, 64}) {
        const int K = 128;
        auto G = rnd(128, K), W = rnd(K, N);
        std::vector<float> ref((size_t)128 * N, 0.f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += G[m * K + k] * W[k * N + n]; ref[m * N + n] = s; }
        std::vector<uint8_t> ia, ib; store_mat(ia, G, 128, K); store_mat(ib, W, K, N);
        Case c{}; c.idesc = idesc(128, N, 0, 1); c.ksteps = K / 16; c.n = N;
        c.a_kstep = 2 * 128 * 16; c.b_kstep = 16 * 16;
        char nm[64];
        snprintf(nm, sizeof nm, "K-major x MN-major N=%d (expected)", N);
        c.a_lbo = 128 * 16; c.a_sbo = 128; c.b_lbo = 128; c.b_sbo = K * 16;
        const bool ok = run(nm, ia, ib, c, ref);
        snprintf(nm, sizeof nm, "K-major x MN-major N=%d (B lbo/sbo swapped)", N);
        c.b_lbo = K * 16; c.b_sbo = 128;
        run(nm, ia, ib, c, ref);
        if (!ok) ++fails;
    }
    printf("umma_probe: %s\n", fails ? "SOME EXPECTED LAYOUTS MISMATCH" : "all expected layouts OK");
    return fails ? 1 : 0;
}
