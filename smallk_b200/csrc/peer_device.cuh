// smallk_b200 — device-side pieces of the NVLink peer-memory exchange (peer.cu), shared with the GEMM whose epilogue stores
// its tiles straight into the owners' receive slots (gemm_f64.cu). Internal header.
#pragma once

#include "common.cuh"
#include "peer.h"

namespace smk {

constexpr unsigned long long kSpinTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

// thread `t` < nranks waits until sender t has published `epoch`; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long epoch)
{
    if (ld_acquire_sys(flag) >= epoch) return true;
    const unsigned long long t0 = global_ns();
    for (;;)
    {
        for (int i = 0; i < 64; ++i)
            if (ld_acquire_sys(flag) >= epoch) return true;
        if (global_ns() - t0 > kSpinTimeoutNs) return false;
        __nanosleep(64);
    }
}

// All threads of the CTA have issued their peer stores. Returns true (to every thread) in the CTA that arrives last.
__device__ __forceinline__ bool stores_done_last_cta(unsigned int* ticket, bool* s_flag)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned int t = atomicAdd(ticket, 1u);
        *s_flag = (t == gridDim.x - 1);
        if (*s_flag) *ticket = 0u;
    }
    __syncthreads();
    return *s_flag;
}

__device__ __forceinline__ unsigned long long* flag_ptr(const PeerTable& t, int receiver, int cls, int sender)
{
    return reinterpret_cast<unsigned long long*>(t.base[receiver] + kPeerFlagOffset) + cls * kPeerMaxRanks + sender;
}


} // namespace smk
