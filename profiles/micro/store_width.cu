// Micro-benchmark behind the SoA packing decision: the same bytes streamed out of (and into) many arrays with 8-byte and
// with 16-byte accesses per thread.  nvcc -arch=sm_100a -O3 store_width.cu -o store_width && ./store_width
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void copy8(const double* const* __restrict__ in, double* const* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v[W];
#pragma unroll
        for (int k = 0; k < W; ++k) v[k] = __ldcs(in[k] + i);
#pragma unroll
        for (int k = 0; k < W; ++k) __stcs(out[k] + i, v[k] * 1.0000001);
    }
}
template <int W>
__global__ void copy16(const double2* const* __restrict__ in, double2* const* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v[W];
#pragma unroll
        for (int k = 0; k < W; ++k) v[k] = __ldcs(in[k] + i);
#pragma unroll
        for (int k = 0; k < W; ++k) { v[k].x *= 1.0000001; v[k].y *= 1.0000001; __stcs(out[k] + i, v[k]); }
    }
}
int main() {
    const size_t n = 10000000;
    const int W8 = 12, W16 = 6;
    double *ibuf, *obuf;
    cudaMalloc(&ibuf, n * 8 * W8 + 4096 * W8);
    cudaMalloc(&obuf, n * 8 * W8 + 4096 * W8);
    cudaMemset(ibuf, 0, n * 8 * W8);
    double* hi[W8]; double* ho[W8];
    for (int k = 0; k < W8; ++k) { hi[k] = ibuf + k * (n + 32); ho[k] = obuf + k * (n + 32); }
    double **di, **dout;
    cudaMalloc(&di, sizeof hi); cudaMalloc(&dout, sizeof ho);
    cudaMemcpy(di, hi, sizeof hi, cudaMemcpyHostToDevice); cudaMemcpy(dout, ho, sizeof ho, cudaMemcpyHostToDevice);
    double2* hi2[W16]; double2* ho2[W16];
    for (int k = 0; k < W16; ++k) { hi2[k] = (double2*)(ibuf + 2 * k * (n + 32)); ho2[k] = (double2*)(obuf + 2 * k * (n + 32)); }
    double2 **di2, **do2;
    cudaMalloc(&di2, sizeof hi2); cudaMalloc(&do2, sizeof ho2);
    cudaMemcpy(di2, hi2, sizeof hi2, cudaMemcpyHostToDevice); cudaMemcpy(do2, ho2, sizeof ho2, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
        for (int rep = 0; rep < 3; ++rep) {
            float m8, m16;
            cudaEventRecord(a); copy8<W8><<<blocks, 256>>>(di, dout, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&m8, a, b);
            cudaEventRecord(a); copy16<W16><<<blocks, 256>>>(di2, do2, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&m16, a, b);
            if (rep == 2) printf("blocks %4d: 12 x 8B arrays %.3f ms (%.0f GB/s)   6 x 16B arrays %.3f ms (%.0f GB/s)\n", blocks, m8,
                                 2.0 * n * 96 / m8 / 1e6, m16, 2.0 * n * 96 / m16 / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
