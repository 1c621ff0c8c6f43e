// issue_probe.cu -- does a packed FP32 instruction (FFMA2) hold the ISSUE port for two cycles, or only the FP32 pipe?
// Streams of FFMA2 (or FFMA) interleaved with independent integer / PRMT / shared-memory instructions; if the other
// instructions issue in the second cycle of an FFMA2, a 1:1 mix runs at ~2 warp-instructions per 2 clocks per scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe/issue_probe tools/probe/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
// FP: 0 none, 1 FFMA, 2 FFMA2.  OTHER: 0 none, 1 LOP3 (xor), 2 PRMT, 3 LDS.32, 4 IMAD, 5 SHFL.  NO = other instructions per FP one
template <int FP, int OTHER, int NO> __global__ void __launch_bounds__(256) k_mix(float *out, int iters, float w, unsigned k) {
    __shared__ unsigned sm[256 * 4];
    float a[8]; unsigned long long p[8]; unsigned x[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 1.f); x[i] = threadIdx.x * 7 + i; }
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i;
    __syncthreads();
    const unsigned long long ww = pk(w, w * 0.5f), cc = pk(0.25f, 0.125f);
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (FP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(w), "f"(a[(i + 1) & 7]));
            if (FP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ww), "l"(cc));
#pragma unroll
            for (int j = 0; j < NO; j++) {
                if (OTHER == 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(k));
                if (OTHER == 2) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[i]) : "r"(k));
                if (OTHER == 3) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(sa + (unsigned)(i * 1024 % 4096)));
                if (OTHER == 4) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(k));
                if (OTHER == 5) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(x[i]));
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) { s += a[i] + (float)x[i]; float u, v; asm("mov.b64 {%0, %1}, %2;" : "=f"(u), "=f"(v) : "l"(p[i])); s += u + v; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int FP, int OTHER, int NO> int run(const char *what, float *d, cudaEvent_t e0, cudaEvent_t e1, int sms) {
    const int iters = 8000;
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k_mix<FP, OTHER, NO><<<sms * 8, 256>>>(d, iters, 0.999f, 0x01020304u);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
    }
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = 1.965;
    const double fp = FP ? (double)sms * 8 * 8 * iters * 8 : 0, oth = (double)sms * 8 * 8 * iters * 8 * NO * (OTHER ? 1 : 0);
    const double clk = ms * 1e-3 * ghz * 1e9;
    printf("%-34s %.3f ms  fp %.2f + other %.2f = %.2f warp-inst/clk/SM   (cycles per FP instruction per scheduler: %.2f)\n", what, ms,
           fp / clk / sms, oth / clk / sms, (fp + oth) / clk / sms, FP ? clk * sms * 4 / fp / 1.0 * 1.0 / 1.0 : 0.0);
    return 0;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d; CK(cudaMalloc(&d, (size_t)sms * 8 * 256 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    run<1, 0, 0>("FFMA alone", d, e0, e1, sms);
    run<2, 0, 0>("FFMA2 alone", d, e0, e1, sms);
    run<0, 1, 1>("LOP3 alone", d, e0, e1, sms);
    run<0, 2, 1>("PRMT alone", d, e0, e1, sms);
    run<0, 3, 1>("LDS.32 alone", d, e0, e1, sms);
    run<0, 4, 1>("IMAD alone", d, e0, e1, sms);
    run<0, 5, 1>("SHFL alone", d, e0, e1, sms);
    run<1, 1, 1>("FFMA : LOP3 1:1", d, e0, e1, sms);
    run<2, 1, 1>("FFMA2 : LOP3 1:1", d, e0, e1, sms);
    run<2, 1, 2>("FFMA2 : LOP3 1:2", d, e0, e1, sms);
    run<2, 2, 1>("FFMA2 : PRMT 1:1", d, e0, e1, sms);
    run<2, 3, 1>("FFMA2 : LDS 1:1", d, e0, e1, sms);
    run<2, 4, 1>("FFMA2 : IMAD 1:1", d, e0, e1, sms);
    run<1, 4, 1>("FFMA : IMAD 1:1", d, e0, e1, sms);
    run<2, 5, 1>("FFMA2 : SHFL 1:1", d, e0, e1, sms);
    return 0;
}
