// Stand-alone probe (GPU box): TMA 2-D loads with the descriptor in kernel-parameter space vs global memory,
// negative start coordinates, u16 elements; FFMA vs FFMA2 throughput.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void load_box(const void *tmap, int x, int y, unsigned char *smem, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = smem_u32(bar), d = smem_u32(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                     "l"(tmap), "r"(x), "r"(y), "r"(b)
                     : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b) : "memory");
}
__global__ void k_param(const __grid_constant__ CUtensorMap tm, int x, int y, unsigned char *out, int bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    load_box(&tm, x, y, sm, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
__global__ void k_global(const CUtensorMap *tm, int x, int y, unsigned char *out, int bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    load_box(tm, x, y, sm, bytes, &bar);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE> __global__ void k_fma(float *out, int iters, float w) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 1.f); }
    const unsigned long long ww = pk(w, w * 0.5f), cc = pk(0.25f, 0.125f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i] = fmaf(a[i], w, 0.25f); }
            else if (MODE == 3) { a[i] = fmaf(a[i], w, a[(i + 1) & 7]); }
            else if (MODE == 1) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ww), "l"(cc)); }
            else { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc)); }
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) { s += a[i]; float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static int check(const std::vector<unsigned char> &img, int W, int H, int pitch, int x0, int y0, int bw_bytes, int bh, const std::vector<unsigned char> &got, const char *what) {
    int bad = 0;
    for (int r = 0; r < bh; r++) for (int c = 0; c < bw_bytes; c++) {
        int xx = x0 + c, yy = y0 + r;
        unsigned char e = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? img[(size_t)yy * pitch + xx] : 0;
        if (got[r * bw_bytes + c] != e) bad++;
    }
    printf("%s: %s (%d mismatches)\n", what, bad ? "FAIL" : "ok", bad);
    return bad;
}
int main(int argc, char **argv) {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    enc_fn enc = (enc_fn)fp;
    const int W = 3840, H = 2160, pitch = 3840;
    std::vector<unsigned char> img((size_t)pitch * H);
    for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d_img, *d_out; CK(cudaMalloc(&d_img, img.size())); CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, 65536));
    auto run = [&](int elem, int boxw, int boxh, int x, int y, bool global, const char *what) {
        CUtensorMap tm; memset(&tm, 0, sizeof(tm));
        cuuint64_t dims[2] = {(cuuint64_t)(W / elem), (cuuint64_t)H}; cuuint64_t str[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh}, es[2] = {1, 1};
        CUresult r = enc(&tm, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, str, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", what, (int)r); return; }
        const int bytes = boxw * elem * boxh;
        CK(cudaMemset(d_out, 0xEE, 65536));
        if (global) {
            CUtensorMap *d_tm; CK(cudaMalloc(&d_tm, sizeof(tm))); CK(cudaMemcpy(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
            k_global<<<1, 128, bytes>>>(d_tm, x, y, d_out, bytes);
        } else k_param<<<1, 128, bytes>>>(tm, x, y, d_out, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: kernel error %s\n", what, cudaGetErrorString(e)); exit(2); }
        std::vector<unsigned char> got(bytes); CK(cudaMemcpy(got.data(), d_out, bytes, cudaMemcpyDeviceToHost));
        check(img, W, H, pitch, x * elem, y, boxw * elem, boxh, got, what);
    };
    if (argc >= 6) {   // tma_probe elem boxw boxh x y
        run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), true, "case");
        return 0;
    }
    run(1, 256, 32, 512, 100, false, "param u8 256x32 interior");
    run(1, 256, 32, 512, 100, true, "global u8 256x32 interior");
    run(1, 256, 32, -16, -3, true, "global u8 256x32 negative 16-byte aligned start");
    run(1, 256, 32, 3696, 2150, true, "global u8 256x32 right/bottom edge");
    run(2, 136, 18, -8, -1, true, "global u16 136x18 negative start (16-byte aligned)");
    run(1, 160, 18, 1904, 5, true, "global u8 160x18");
    // NOTE: a box whose first byte is not 16-byte aligned (x * elem % 16 != 0) raises "illegal instruction": run the
    // probe with arguments `elem boxw boxh x y` to see it (it poisons the context, so not part of the default run)
    // FFMA vs FFMA2 throughput
    float *d_f; CK(cudaMalloc(&d_f, 148 * 8 * 256 * 4 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k_fma<0><<<148 * 8, 256>>>(d_f, iters, 0.999f); else if (mode == 1) k_fma<1><<<148 * 8, 256>>>(d_f, iters, 0.999f); else if (mode == 2) k_fma<2><<<148 * 8, 256>>>(d_f, iters, 0.999f); else k_fma<3><<<148 * 8, 256>>>(d_f, iters, 0.999f);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)148 * 8 * 8 * iters * 8;   // warp-instructions
        printf("%s: %.3f ms, %.2f warp-inst/clk/SM @1.965GHz (%.1f TFLOP/s fp32 eq)\n", mode == 0 ? "FFMA (imm addend)" : mode == 1 ? "FFMA2" : mode == 2 ? "FADD2" : "FFMA (3 registers)", ms,
               inst / (ms * 1e-3) / 1.965e9 / 148, inst * 32 * (mode == 1 ? 4 : 2) / (ms * 1e-3) / 1e12);
    }
    return 0;
}
