// DMMA.8x8x4 throughput against warps per SM and independent accumulator chains per warp (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_sweep dmma_sweep.cu && ./dmma_sweep
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// every second DMMA predicated OFF (warp-uniform predicate the compiler cannot see through): does an issued but
// disabled DMMA hold the tensor pipe?
__device__ __forceinline__ void dmma_pred(double &c0, double &c1, double a, double b, int on) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %4, 0;\n@p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n}\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b), "r"(on));
}
template <int CH>
__global__ void kpred(int iters, double *out, int off_flag) {
    double c[CH][2];
#pragma unroll
    for (int u = 0; u < CH; ++u) { c[u][0] = threadIdx.x * 1e-3 + u; c[u][1] = c[u][0] + 0.5; }
    const double a = 1.0000001 * ((threadIdx.x & 3) == 0 ? 1.0 : 1e-9), b = 0.9999999 * ((threadIdx.x & 3) == 0 ? 1.0 : 1e-9);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < CH; ++u) dmma_pred(c[u][0], c[u][1], a, b, (u & 1) ? off_flag : 1);
    }
    double r = 0.0;
#pragma unroll
    for (int u = 0; u < CH; ++u) r += c[u][0] + c[u][1];
    if (r == 12345.678) out[0] = r;
}
template <int CH, bool LDSFEED>
__global__ void k(int iters, double *out) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 1e-9 * i;
    __syncthreads();
    double c[CH][2];
#pragma unroll
    for (int u = 0; u < CH; ++u) { c[u][0] = threadIdx.x * 1e-3 + u; c[u][1] = c[u][0] + 0.5; }
    double a = 1.0000001 * ((threadIdx.x & 3) == 0 ? 1.0 : 1e-9), b = 0.9999999 * ((threadIdx.x & 3) == 0 ? 1.0 : 1e-9);
    const double2 *p = reinterpret_cast<const double2 *>(sh) + (threadIdx.x & 31);
    for (int i = 0; i < iters; ++i) {
        if (LDSFEED) { const double2 v = p[(i & 7) * 32], w = p[((i + 3) & 7) * 32 + 256]; a = v.x; b = w.y; }
#pragma unroll
        for (int u = 0; u < CH; ++u) dmma(c[u][0], c[u][1], a, b);
    }
    double r = 0.0;
#pragma unroll
    for (int u = 0; u < CH; ++u) r += c[u][0] + c[u][1];
    if (r == 12345.678) out[0] = r;
}
template <int CH, bool L>
void run(int warps_per_sm, int sms, double *out) {
    const int iters = 20000;
    const int block = warps_per_sm >= 8 ? 256 : 32 * warps_per_sm, ctas = warps_per_sm >= 8 ? warps_per_sm / 8 : 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, L><<<sms * ctas, block>>>(200, out);
    cudaEventRecord(e0); k<CH, L><<<sms * ctas, block>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)sms * warps_per_sm * iters * CH * 512.0;
    printf("warps/SM %2d chains %d lds %d: %6.2f TFLOP/s  (%.1f cycles per DMMA per SMSP at 1.965 GHz)\n", warps_per_sm, CH, (int)L, fl / ms / 1e9,
           ms * 1e-3 * 1.965e9 / ((double)warps_per_sm / 4 * iters * CH));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8);
    for (int w : {4, 8, 16, 32, 64}) { run<1, false>(w, p.multiProcessorCount, out); run<2, false>(w, p.multiProcessorCount, out); run<4, false>(w, p.multiProcessorCount, out); run<8, false>(w, p.multiProcessorCount, out); }
    for (int w : {8, 16, 32}) { run<4, true>(w, p.multiProcessorCount, out); run<8, true>(w, p.multiProcessorCount, out); }
    for (int off : {1, 0}) {
        const int iters = 20000, sms = p.multiProcessorCount;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        kpred<4><<<sms * 2, 256>>>(200, out, off);
        cudaEventRecord(e0); kpred<4><<<sms * 2, 256>>>(iters, out, off); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("16 warps/SM, 4 chains, odd chains predicated %s: %.3f ms (%.1f cycles per ISSUED DMMA per SMSP)\n", off ? "ON" : "OFF", ms,
               ms * 1e-3 * 1.965e9 / (4.0 * iters * 4));
    }
    return 0;
}
