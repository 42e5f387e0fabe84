// smallk_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

namespace smk {

constexpr int kWarp = 32;

// ---- error plumbing -------------------------------------------------------
struct CudaError { cudaError_t code; const char* file; int line; };

#define SMK_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) throw ::smk::CudaError{_e, __FILE__, __LINE__}; \
    } while (0)

// every kernel launch of this library goes through here, so it can be counted (bench.py: gpu_launches); atomic because two
// contexts may launch from two host threads (the hierclust driver's worker thread sorts on its own context)
inline std::atomic<long long>& launch_counter() { static std::atomic<long long> n{0}; return n; }
#define SMK_LAUNCH_CHECK()                 \
    do {                                   \
        ++::smk::launch_counter();         \
        SMK_CUDA(cudaGetLastError());      \
    } while (0)

// ---- cp.async (LDGSTS) ----------------------------------------------------
// 16-byte and 8-byte asynchronous global->shared copies with zero fill of the
// bytes beyond src_bytes (src_bytes may be 0: nothing is read, all zeros).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- FP64 tensor core: DMMA.8x8x4 ----------------------------------------
// D(8x8) += A(8x4, row) * B(4x8, col). Lane l holds A[l>>2][l&3], B[l&3][l>>2],
// C[l>>2][2*(l&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- warp reductions -------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

} // namespace smk
