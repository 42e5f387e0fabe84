// Measures what random gathers of whole k-vectors sustain on this GPU — the access pattern of a gather SpMM
// (csrc/spmm.cu): a warp (or a group of lanes) reads `vec_bytes` contiguous bytes at a pseudo-random vector index of
// a table of `table_MB`, `U` independent gathers in flight per lane, and adds what it read. No index stream, no
// arithmetic to speak of: the number printed is the rate at which the memory system delivers gathered operand bytes
// to the SMs — from L2 when the table fits, from HBM through L2 when it does not. It is the denominator the SpMM
// kernels are held against (bench.py --workload c3: roofline.l2_gather_bound), next to the guide's 6300 B/clk figure.
//
//   l2_gather_peak            prints one line per (table size, vector size): GB/s gathered
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

struct alignas(32) V4 { double x, y, z, w; };

__device__ __forceinline__ V4 ld256(const double* p)
{
    V4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// LPV lanes per vector (vec_bytes = LPV * 32), U gathers in flight per lane
template <int LPV, int U>
__global__ void __launch_bounds__(512) gather_kernel(const double* __restrict__ table, uint32_t nvec, int iters, double* __restrict__ out, uint32_t stride_d)
{
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = gtid / LPV, lane = gtid % LPV;
    double acc = 0.0;
    uint32_t s = group * 0x9e3779b9u + 12345u;
    for (int it = 0; it < iters; ++it)
    {
        V4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            s = mix(s + 0x632be5abu);
            const uint32_t idx = static_cast<uint32_t>((static_cast<uint64_t>(s) * nvec) >> 32);
            v[u] = ld256(table + static_cast<size_t>(idx) * stride_d + lane * 4);
        }
#pragma unroll
        for (int u = U - 1; u >= 0; --u) { acc += v[u].x; acc += v[u].y; acc += v[u].z; acc += v[u].w; }   // one chain, the LAST load consumed first: all U are in flight
    }
    out[gtid] = acc;
}

template <int LPV, int U>
double run(const double* table, size_t table_bytes, int sms, int ctas_per_sm, int threads, double* out, int stride_bytes = 0)
{
    if (stride_bytes == 0) stride_bytes = LPV * 32;
    const uint32_t nvec = static_cast<uint32_t>(table_bytes / stride_bytes);
    const int iters = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
    {
        cudaEventRecord(e0);
        gather_kernel<LPV, U><<<sms * ctas_per_sm, threads>>>(table, nvec, iters, out, static_cast<uint32_t>(stride_bytes / 8));
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = 32.0 * U * iters * static_cast<double>(sms) * ctas_per_sm * threads;
    return bytes / best * 1e-6;    // GB/s
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const size_t big = size_t(1) << 30;
    double* table; cudaMalloc(&table, big);
    cudaMemset(table, 0, big);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
    const size_t sizes_mb[] = {16, 32, 48, 64, 205, 1024};
    for (size_t mb : sizes_mb)
    {
        const size_t bytes = mb << 20;
        // 1 KB vectors (k = 128, a warp per gathered vector), 8 and 4 in flight; 256 B vectors (a 32-row slab, 8 lanes per vector)
        const double a = run<32, 8>(table, bytes, sms, 2, 512, out);
        const double b = run<32, 4>(table, bytes, sms, 4, 512, out);
        const double c = run<8, 8>(table, bytes, sms, 2, 512, out);
        printf("GATHER table=%4zu MB : 1KB-vectors U=8 x32 warps/SM %.0f GB/s | U=4 x64 warps/SM %.0f GB/s | 256B-vectors U=8 %.0f GB/s\n",
               mb, a, b, c);
    }
    // how many bytes in flight the rate needs: warps per SM x gathers in flight per lane (table 48 MB, L2-resident)
    {
        const size_t bytes = size_t(48) << 20;
        printf("INFLIGHT 1KB-vectors, 48 MB table: warps/SM x U -> GB/s :");
        printf(" 8x8 %.0f |", run<32, 8>(table, bytes, sms, 1, 256, out));
        printf(" 16x8 %.0f |", run<32, 8>(table, bytes, sms, 1, 512, out));
        printf(" 16x4 %.0f |", run<32, 4>(table, bytes, sms, 1, 512, out));
        printf(" 32x4 %.0f |", run<32, 4>(table, bytes, sms, 2, 512, out));
        printf(" 32x2 %.0f |", run<32, 2>(table, bytes, sms, 2, 512, out));
        printf(" 64x2 %.0f |", run<32, 2>(table, bytes, sms, 4, 512, out));
        printf(" 64x1 %.0f\n", run<32, 1>(table, bytes, sms, 4, 512, out));
    }
    // the k-slab pattern of spmm_seg_slab_kernel: 256-byte pieces of 1-KB vectors (a 32-row slab of a k = 128 operand of 205 MB)
    {
        const size_t bytes = size_t(205) << 20;
        printf("SLAB 256B pieces at 1 KB stride, 205 MB operand (51 MB touched): 16 warps/SM U=8 %.0f | 32 warps/SM U=8 %.0f | 64 warps/SM U=4 %.0f GB/s\n",
               run<8, 8>(table, bytes, sms, 1, 512, out, 1024), run<8, 8>(table, bytes, sms, 2, 512, out, 1024), run<8, 4>(table, bytes, sms, 4, 512, out, 1024));
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
