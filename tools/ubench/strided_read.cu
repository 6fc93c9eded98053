// What HBM delivers for the access pattern of hessian_grad_dmma_kernel: every CTA streams ITS scenario's Gamma (2N x N doubles,
// column-major, 160 KB at N = 100) in row chunks -- per chunk 100 pieces of 448 B at a stride of 1,600 B -- against the same
// bytes read contiguously.  296 persistent CTAs of 256 threads, 16-byte cp.async into shared memory, nothing else.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_read strided_read.cu && ./strided_read
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void cp16(void *d, const void *s) {
    unsigned a = (unsigned)__cvta_generic_to_shared(d);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(s) : "memory");
}
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int MODE>   // 0: row chunks of KC rows of every column; 1: the same number of bytes per step, contiguous
__global__ void __launch_bounds__(256, 2) k(int S, int N, int KC, const double *G, double *out) {
    extern __shared__ __align__(16) double sm[];
    const int K2 = 2 * N, KH = KC / 2;
    double acc = 0.0;
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        const double *g = G + (size_t)s * K2 * N;
        for (int kc = 0; kc < K2; kc += KC) {
            const int rows = min(KC, K2 - kc), hr = rows / 2;
            __syncthreads();
            for (int p = threadIdx.x; p < N * hr; p += 256) {
                if (MODE == 0) { const int c = p / hr, h = p - c * hr; cp16(sm + 2 * (c * KH + h), g + (size_t)c * K2 + kc + 2 * h); }
                else cp16(sm + 2 * p, g + (size_t)kc * N + 2 * p);
            }
            wait_all();
            __syncthreads();
            acc += sm[threadIdx.x];
        }
    }
    if (acc == 1234.5) out[0] = acc;
}
int main() {
    const int S = 16384, N = 100;
    double *G, *out; cudaMalloc(&G, (size_t)S * 2 * N * N * 8); cudaMalloc(&out, 8); cudaMemset(G, 0, (size_t)S * 2 * N * N * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int KC : {56, 100, 200}) {
        const size_t smem = (size_t)N * KC * 8;
        cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = KC <= 100 ? 296 : 148;
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9;
            for (int r = 0; r < 4; ++r) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, 256, smem>>>(S, N, KC, G, out); else k<1><<<grid, 256, smem>>>(S, N, KC, G, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
            }
            printf("KC %3d (%4d B pieces) %s: %.3f ms  %.0f GB/s  [%s]\n", KC, KC * 8, mode ? "contiguous " : "row chunks ", best,
                   (double)S * 2 * N * N * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
