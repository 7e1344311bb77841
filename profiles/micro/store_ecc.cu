// Which store pattern makes DRAM reads (ECC read-modify-write) appear?  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <typename T, bool CS>
__global__ void st_kernel(T* __restrict__ p, size_t n, int arrays, size_t pitch) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        for (int k = 0; k < arrays; ++k) { if (CS) __stcs(p + k * pitch + i, (T)(i + k)); else p[k * pitch + i] = (T)(i + k); }
}
// stores of one array separated by work on other arrays (as in the fused kernel: hit records mid-way, bundle at the end)
__global__ void st_spread(double* __restrict__ p, size_t n, int arrays, size_t pitch, int spin) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v = (double)i;
        for (int k = 0; k < arrays; ++k) {
            for (int s = 0; s < spin; ++s) v = fma(v, 1.0000001, 1e-9);
            __stcs(p + k * pitch + i, v);
        }
    }
}
int main() {
    const size_t n = 10000000, pitch = n + 64;
    void* buf; cudaMalloc(&buf, pitch * 8 * 40);
    st_kernel<uint8_t, false><<<740, 128>>>((uint8_t*)buf, n, 8, pitch);
    st_kernel<uint32_t, true><<<740, 128>>>((uint32_t*)buf, n, 8, pitch);
    st_kernel<double, true><<<740, 128>>>((double*)buf, n, 8, pitch);
    st_kernel<double, false><<<740, 128>>>((double*)buf, n, 8, pitch);
    st_spread<<<740, 128>>>((double*)buf, n, 8, pitch, 200);
    st_spread<<<740, 128>>>((double*)buf, n, 40, pitch, 40);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
