// Measures what the FP64 tensor pipe (DMMA.8x8x4) and the FP64 FMA pipe sustain on this GPU:
// register-resident, no memory traffic. Prints TFLOP/s. Used once per round to put a number
// next to cuBLAS DGEMM as the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dmma_loop(double* out, int iters, double a0, double b0)
{
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x, b = b0 + threadIdx.x;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dfma_loop(double* out, int iters, double a0, double b0)
{
    double c[ILP];
    for (int i = 0; i < ILP; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2)
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            cudaEventRecord(e0);
            dmma_loop<16><<<sms, warps * 32>>>(out, iters, 1.0, 2.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256 * 16 * (double)iters * warps * sms;
            if (rep) printf("DMMA.8x8x4  warps/SM=%2d ILP=16 : %.2f TFLOP/s (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
    }
    for (int warps = 8; warps <= 32; warps *= 2)
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            cudaEventRecord(e0);
            dfma_loop<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 8 * (double)iters * warps * sms;
            if (rep) printf("DFMA        warps/SM=%2d ILP=8  : %.2f TFLOP/s (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
