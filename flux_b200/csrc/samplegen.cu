// samplegen.cu — MasterSampleSets::new on the device.
//
// Restates fluxcore/src/sampling.rs:13-40 and samplers/src/lib.rs:46-182:
//   pixel_sets[s] = correlated multi-jittered grid (CMJ)        sampling.rs:16-17
//   disc_sets[s]  = to_poisson_disc(CMJ)                        sampling.rs:19-21
//   hemi_sets[s][d] = to_hemisphere(multi-jittered grid, e=0)   sampling.rs:23-29
//   per-row permutation of set indices                           sampling.rs:35-40
//
// The reference draws from an unseeded IsaacRng (samplers/src/lib.rs:27-33),
// so its streams cannot be reproduced.  This generator uses a
// counter-based PRNG (splitmix64 finaliser keyed per stream, DESIGN.md
// "PRNG") so that every grid, row of a grid and image row is an independent
// stream that one CTA / one thread can produce without communication.  The
// oracle implements the same streams on the CPU, so permutations and unit-
// square coordinates agree bit-for-bit; disc and hemisphere maps agree to the
// last ulps of sin/cos.
//
// One CTA per (set, grid).  Permutations live in shared memory as u16.
#include "flux_kernels.cuh"
#include "flux_shade.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t k = mix64(seed + 0x9E3779B97F4A7C15ull);
    k = mix64(k + a);
    k = mix64(k + b);
    k = mix64(k + c);
    k = mix64(k + d);
    return k;
}
__device__ __forceinline__ uint64_t rnd(uint64_t key, uint64_t ctr) {
    return mix64(key + (ctr + 1) * 0x9E3779B97F4A7C15ull);
}
__device__ __forceinline__ double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ uint32_t below(uint64_t x, uint32_t n) { return (uint32_t)__umul64hi(x, (uint64_t)n); }

enum { P_JITTER = 0, P_PERM_Y = 1, P_PERM_X = 2, P_ROW = 3 };

// Rng::shuffle (rand 0.5.5): for i = len-1 down to 1: swap(i, gen_range(0, i+1))
template <class T> __device__ void fisher_yates(T *v, uint32_t len, uint64_t key) {
    for (uint32_t i = len; i >= 2;) {
        i -= 1;
        uint32_t j = below(rnd(key, i), i + 1);
        T tmp = v[i];
        v[i] = v[j];
        v[j] = tmp;
    }
}

// to_poisson_disc (one point), samplers/src/lib.rs:146-180
__device__ __forceinline__ double2 to_disc(double px, double py) {
    double spx = 2.0 * px - 1.0;
    double spy = 2.0 * py - 1.0;
    double phi, r;
    if (spx > -spy) {
        if (spx > spy) {
            r = spx;
            phi = spy / spx;
        } else {
            r = spy;
            phi = 2.0 - spx / spy;
        }
    } else {
        if (spx < spy) {
            r = -spx;
            phi = 4.0 + spy / spx;
        } else {
            r = -spy;
            if (spy != 0.0)
                phi = 6.0 - spx / spy;
            else
                phi = 0.0;
        }
    }
    phi *= FLUX_PI / 4.0;
    double s, c;
    sincos(phi, &s, &c);
    return make_double2(r * c, r * s);
}

// to_unit_hemi(p, 0.0), samplers/src/lib.rs:133-142: powf(1-y, 1/(0+1)) == 1-y exactly.
__device__ __forceinline__ V3 to_unit_hemi_e0(double px, double py) {
    double sin_phi, cos_phi;
    sincos((2.0 * FLUX_PI) * px, &sin_phi, &cos_phi);
    double cos_theta = 1.0 - py;
    double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    return normalize3(mk3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
}

// grid 0 = pixel (CMJ), 1 = disc (CMJ), 2+d = hemisphere depth d (MJ)
__global__ void __launch_bounds__(256) generate_samples_kernel(uint64_t seed, uint32_t root, uint32_t max_depth,
                                                               double2 *__restrict__ pixel, double2 *__restrict__ disc,
                                                               double *__restrict__ hemi) {
    extern __shared__ uint16_t perm[];  // py[line][k] then px[line][k]; CMJ uses line 0 only
    const uint32_t grids = 2 + max_depth;
    const uint32_t set = blockIdx.x / grids;
    const uint32_t grid = blockIdx.x % grids;
    const bool correlated = grid < 2;
    const uint32_t lines = correlated ? 1 : root;
    uint16_t *py = perm;
    uint16_t *px = perm + (size_t)lines * root;
    for (uint32_t k = threadIdx.x; k < lines * root; k += blockDim.x) {
        py[k] = (uint16_t)(k % root);
        px[k] = (uint16_t)(k % root);
    }
    __syncthreads();
    // shuffle_y / shuffle_x permutations, samplers/src/lib.rs:92-126 (one per line for MJ,
    // one shared for CMJ lib.rs:78-82)
    for (uint32_t w = threadIdx.x; w < 2 * lines; w += blockDim.x) {
        uint32_t line = w >> 1;
        if (w & 1)
            fisher_yates(px + (size_t)line * root, root, stream_key(seed, set, grid, P_PERM_X, line));
        else
            fisher_yates(py + (size_t)line * root, root, stream_key(seed, set, grid, P_PERM_Y, line));
    }
    __syncthreads();
    const uint64_t kj = stream_key(seed, set, grid, P_JITTER, 0);
    const double r_float = (double)root;
    const double r2 = (double)((unsigned long long)root * root);
    const uint32_t n = root * root;
    for (uint32_t idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const uint32_t i = idx / root, j = idx - i * root;
        const uint32_t xi = px[correlated ? i : j * root + i];  // pix_j(i)
        const uint32_t yj = py[correlated ? j : i * root + j];  // piy_i(j)
        // grid_multi_jittered_base, samplers/src/lib.rs:46-62: x from cell (xi, j), y from cell (i, yj)
        const double a = u01(rnd(kj, 2ull * ((uint64_t)xi * root + j)));
        const double b = u01(rnd(kj, 2ull * ((uint64_t)i * root + yj) + 1));
        const double x = ((double)xi / r_float) + ((double)(root - 1 - j) + a) / r2;
        const double y = ((double)yj / r_float) + ((double)(root - 1 - i) + b) / r2;
        if (grid == 0) {
            pixel[(size_t)set * n + idx] = make_double2(x, y);
        } else if (grid == 1) {
            disc[(size_t)set * n + idx] = to_disc(x, y);
        } else {
            V3 h = to_unit_hemi_e0(x, y);
            double *o = hemi + (((size_t)set * max_depth + (grid - 2)) * n + idx) * 3;
            o[0] = h.x;
            o[1] = h.y;
            o[2] = h.z;
        }
    }
}

// One CTA per image row: shuffle_indices (sampling.rs:35-40), idx[row][col] = perm[col % num_sets]
__global__ void __launch_bounds__(256) generate_set_index_kernel(uint64_t seed, uint32_t W, uint32_t num_sets,
                                                                 uint32_t *__restrict__ idx) {
    extern __shared__ uint32_t rperm[];
    const uint32_t row = blockIdx.x;
    for (uint32_t k = threadIdx.x; k < num_sets; k += blockDim.x) rperm[k] = k;
    __syncthreads();
    if (threadIdx.x == 0) fisher_yates(rperm, num_sets, stream_key(seed, 0xFFFFFFFFull, (uint64_t)row, P_ROW, 0));
    __syncthreads();
    for (uint32_t col = threadIdx.x; col < W; col += blockDim.x) idx[(size_t)row * W + col] = rperm[col % num_sets];
}

// Glossy lobe table: ghemi[set][k][i] = to_unit_hemi(pixel_sets[set][i], exp_k) (samplers/src/lib.rs:133-142 as
// called at brdf.rs:64).  The same device function the render kernels would call inline, so the table holds
// bit-identical values; it removes sincos + pow + sqrt + 3 divisions from every glossy bounce.
__global__ void __launch_bounds__(256) build_glossy_table_kernel(const double2 *__restrict__ pixel, uint32_t n,
                                                                 uint32_t num_sets, uint32_t gk,
                                                                 const double *__restrict__ inv_e1,
                                                                 double *__restrict__ ghemi) {
    const size_t total = (size_t)num_sets * gk * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(t % n);
        const uint32_t k = (uint32_t)((t / n) % gk);
        const size_t set = t / ((size_t)n * gk);
        const double2 p = pixel[set * n + i];
        const V3 h = to_unit_hemi_dev(p.x, p.y, inv_e1[k]);
        ghemi[3 * t + 0] = h.x;
        ghemi[3 * t + 1] = h.y;
        ghemi[3 * t + 2] = h.z;
    }
}

}  // namespace

void launch_build_glossy_table(const double2 *pixel, uint32_t n, uint32_t num_sets, uint32_t gk, const double *inv_e1,
                               double *ghemi, int sm_count, cudaStream_t stream) {
    build_glossy_table_kernel<<<sm_count * 8, 256, 0, stream>>>(pixel, n, num_sets, gk, inv_e1, ghemi);
}

void launch_generate_samples(uint64_t seed, uint32_t root, uint32_t max_depth, uint32_t num_sets, double2 *pixel,
                             double2 *disc, double *hemi, cudaStream_t stream) {
    const size_t smem = (size_t)2 * root * root * sizeof(uint16_t);
    cudaFuncSetAttribute(generate_samples_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    generate_samples_kernel<<<num_sets * (2 + max_depth), 256, smem, stream>>>(seed, root, max_depth, pixel, disc, hemi);
}

void launch_generate_set_index(uint64_t seed, uint32_t H, uint32_t W, uint32_t num_sets, uint32_t *idx,
                               cudaStream_t stream) {
    const size_t smem = (size_t)num_sets * sizeof(uint32_t);
    cudaFuncSetAttribute(generate_set_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    generate_set_index_kernel<<<H, 256, smem, stream>>>(seed, W, num_sets, idx);
}

// ---- FP64 issue-rate microbenchmark (roofline denominator) -----
// Eight independent dependent-chains of alternating DADD / DMUL per thread
// (no FMA: -fmad=false), 8 warps per SM sub-partition.
__global__ void __launch_bounds__(1024) fp64_peak_kernel(int iters, double *sink) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = a0 * m; a1 = a1 * m; a2 = a2 * m; a3 = a3 * m;
            a4 = a4 * m; a5 = a5 * m; a6 = a6 * m; a7 = a7 * m;
            a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c;
            a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;  // never true: keeps the chains alive
}

double launch_fp64_peak(int sm_count, int iters, double *sink, cudaStream_t stream) {
    const int threads = 1024, blocks = sm_count * 2;
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
    return (double)blocks * threads * (double)iters * 8.0 * 16.0;  // thread-level FP64 instructions
}
