// Write-only HBM bandwidth on B200: what a kernel that only streams results out can reach, by store flavour and CTA shape.
// (MEASURED_PEAKS.json's 6552 GB/s is a copy: half of its bytes are reads.)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void wr(float4* p, size_t n, float v) {
    const float4 x = make_float4(v, v + 1, v + 2, v + 3);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) p[i] = x;
        else if (MODE == 1) __stcs(p + i, x);
        else __stwt(p + i, x);
    }
}
// every CTA writes contiguous 32 KB pieces (like the interpolation kernel's tiles)
__global__ void wr_tiles(float4* p, size_t n, float v) {
    const float4 x = make_float4(v, v + 1, v + 2, v + 3);
    const size_t per = 2048;   // float4 per tile
    for (size_t t = blockIdx.x; t * per < n; t += gridDim.x)
        for (size_t i = threadIdx.x; i < per; i += blockDim.x) __stcs(p + t * per + i, x);
}
__global__ void cp(const float4* a, float4* b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
    const size_t bytes = (size_t)8 << 30, n = bytes / 16;
    float4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto f, double nbytes) {
        float best = 1e9;
        for (int r = 0; r < 6; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
        printf("%-28s %8.3f ms  %8.1f GB/s\n", name, best, nbytes / best / 1e6);
    };
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32})
        for (int t : {128, 256, 512}) {
            char nm[64];
            snprintf(nm, 64, "st   grid %5d x %3d", g, t); run(nm, [&] { wr<0><<<g, t>>>(a, n, 1.f); }, (double)bytes);
            snprintf(nm, 64, "stcs grid %5d x %3d", g, t); run(nm, [&] { wr<1><<<g, t>>>(a, n, 1.f); }, (double)bytes);
        }
    run("stwt 148*16 x 256", [&] { wr<2><<<148 * 16, 256>>>(a, n, 1.f); }, (double)bytes);
    run("tiles 32KB 148*6 x 128", [&] { wr_tiles<<<148 * 6, 128>>>(a, n, 1.f); }, (double)bytes);
    run("tiles 32KB 148*12 x 128", [&] { wr_tiles<<<148 * 12, 128>>>(a, n, 1.f); }, (double)bytes);
    run("copy 148*16 x 256 (r+w)", [&] { cp<<<148 * 16, 256>>>(a, b, n); }, 2.0 * bytes);
    run("cudaMemset", [&] { cudaMemsetAsync(a, 1, bytes); }, (double)bytes);
    run("cudaMemcpy d2d (r+w)", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 2.0 * bytes);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
