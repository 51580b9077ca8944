// sm_100a layout kernels (K6 of SURVEY §2.4): host column layout [body][width] <-> device SoA planes,
// and the FP64 throughput probe.
#include <algorithm>

#include "sixdof_internal.h"
#include "sixdof_launch.h"

namespace b200 {

// ================================================================== layout kernels (K6)

static constexpr int kTile = 256;

__global__ void __launch_bounds__(kTile) aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa,
                                                           uint64_t n_bodies, uint32_t width, uint64_t ld)
{
    extern __shared__ double tile[]; // kTile * (width | 1)
    const uint32_t pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    const double *src = aos + base * width;
    for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) tile[(i / width) * pitch + (i % width)] = src[i];
    __syncthreads();
    if (threadIdx.x < nb)
        for (uint32_t k = 0; k < width; ++k) soa[(uint64_t)k * ld + base + threadIdx.x] = tile[threadIdx.x * pitch + k];
}

// gridDim.y = samples; a sample's planes start at soa + y*width*ld, its rows at aos + y*n_bodies*width
__global__ void __launch_bounds__(kTile) soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos,
                                                           uint64_t n_bodies, uint32_t width, uint64_t ld)
{
    extern __shared__ double tile[];
    const uint32_t pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    const double *s = soa + (uint64_t)blockIdx.y * width * ld;
    double *dst = aos + (uint64_t)blockIdx.y * n_bodies * width + base * width;
    if (threadIdx.x < nb)
        for (uint32_t k = 0; k < width; ++k) tile[threadIdx.x * pitch + k] = s[(uint64_t)k * ld + base + threadIdx.x];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) dst[i] = tile[(i / width) * pitch + (i % width)];
}

// All columns of a small batch in one launch: blockIdx.y selects the column, the AoS side of every
// column lives in one packed staging buffer (one PCIe transfer per direction per invoke_batch).
__global__ void __launch_bounds__(kTile) multi_transpose_kernel(const __grid_constant__ MultiColumns mc, uint64_t n_bodies, uint64_t ld,
                                                                int to_soa)
{
    extern __shared__ double tile[];
    const MultiColumns::Col c = mc.col[blockIdx.y];
    const uint32_t width = c.width, pitch = width | 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kTile;
    if (base >= n_bodies) return;
    const uint32_t nb = (uint32_t)min((uint64_t)kTile, n_bodies - base);
    double *aos = mc.packed + c.aos_offset + base * width;
    if (to_soa) {
        for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) tile[(i / width) * pitch + (i % width)] = aos[i];
        __syncthreads();
        if (threadIdx.x < nb)
            for (uint32_t k = 0; k < width; ++k) c.soa[(uint64_t)k * ld + base + threadIdx.x] = tile[threadIdx.x * pitch + k];
    } else {
        if (threadIdx.x < nb)
            for (uint32_t k = 0; k < width; ++k) tile[threadIdx.x * pitch + k] = c.soa[(uint64_t)k * ld + base + threadIdx.x];
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb * width; i += kTile) aos[i] = tile[(i / width) * pitch + (i % width)];
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) probe_fp64_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(uint64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ================================================================== launchers

cudaError_t launch_aos_to_soa(const double *aos, double *soa, uint64_t n_bodies, uint32_t width, uint64_t ld, cudaStream_t s)
{
    if (n_bodies == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_bodies + kTile - 1) / kTile);
    aos_to_soa_kernel<<<grid, kTile, kTile * (width | 1u) * sizeof(double), s>>>(aos, soa, n_bodies, width, ld);
    return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(const double *soa, double *aos, uint64_t n_bodies, uint32_t width, uint64_t ld, cudaStream_t s)
{
    if (n_bodies == 0) return cudaSuccess;
    const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), 1);
    soa_to_aos_kernel<<<grid, kTile, kTile * (width | 1u) * sizeof(double), s>>>(soa, aos, n_bodies, width, ld);
    return cudaGetLastError();
}

cudaError_t launch_traj_to_aos(const double *traj, double *aos, uint64_t n_samples, uint64_t n_bodies, uint64_t ld,
                               uint32_t width, cudaStream_t s)
{
    if (n_bodies == 0 || n_samples == 0) return cudaSuccess;
    const size_t smem = (size_t)kTile * (width | 1u) * sizeof(double); // 25 planes: 52 KB, above the 48 KB default
    cudaError_t e = ensure_dynamic_smem(soa_to_aos_kernel, smem);
    if (e != cudaSuccess) return e;
    for (uint64_t s0 = 0; s0 < n_samples; s0 += 32768) {
        const unsigned ny = (unsigned)min((uint64_t)32768, n_samples - s0);
        const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), ny);
        soa_to_aos_kernel<<<grid, kTile, smem, s>>>(traj + s0 * width * ld, aos + s0 * n_bodies * width, n_bodies, width, ld);
    }
    return cudaGetLastError();
}

cudaError_t launch_multi_transpose(const MultiColumns &mc, uint64_t n_bodies, uint64_t ld, bool to_soa, cudaStream_t s)
{
    if (n_bodies == 0 || mc.n == 0) return cudaSuccess;
    uint32_t wmax = 1;
    for (uint32_t i = 0; i < mc.n; ++i) wmax = std::max(wmax, mc.col[i].width);
    const dim3 grid((unsigned)((n_bodies + kTile - 1) / kTile), mc.n);
    multi_transpose_kernel<<<grid, kTile, kTile * (wmax | 1u) * sizeof(double), s>>>(mc, n_bodies, ld, to_soa ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_probe_fp64(double *out, int iters, int blocks, cudaStream_t s)
{
    probe_fp64_kernel<<<blocks, 256, 0, s>>>(out, iters);
    return cudaGetLastError();
}


} // namespace b200
