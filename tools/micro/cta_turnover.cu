// How fast can the GPU turn over small CTAs?  nvcc -arch=sm_100a -O3 -o /tmp/cta_turnover tools/micro/cta_turnover.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
template <int MINB> __global__ void __launch_bounds__(256, MINB) k_empty(const unsigned char *f, float4 *img, int mode, const unsigned *range) {
    if (mode == 0) return;
    if (mode == 4) { // range check through L1 (same two words for every CTA)
        const unsigned lo = __ldg(range), hi = __ldg(range + 1);
        if (blockIdx.x < lo || blockIdx.x >= hi) return;
    }
    const unsigned char c = f[blockIdx.x];
    if (mode == 1) { if (c) img[0] = make_float4(1, 1, 1, 1); return; }
    if (c == 0) { // background: 3 KB
        if (mode == 2) { if (threadIdx.x < 192) __stcs(img + (size_t)blockIdx.x * 192 + threadIdx.x, make_float4(0, 0, 0, 0)); }
        else if (mode == 3) { if (threadIdx.x < 32) for (int i = 0; i < 6; i++) __stcs(img + (size_t)blockIdx.x * 192 + i * 32 + threadIdx.x, make_float4(0, 0, 0, 0)); }
        return;
    }
    img[0] = make_float4(1, 1, 1, 1);
}
__global__ void k_fill(float4 *img, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(img + i, make_float4(0, 0, 0, 0));
}
int main() {
    const int n = 8100;
    unsigned char *f; float4 *img; unsigned *range; float *flush;
    CK(cudaMalloc(&f, n)); CK(cudaMemset(f, 0, n));
    CK(cudaMalloc(&img, (size_t)n * 192 * 16)); CK(cudaMalloc(&range, 8));
    unsigned hr[2] = {2696, 5404}; CK(cudaMemcpy(range, hr, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&flush, 256u << 20));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        float best = 1e9, sum = 0; const int reps = 30;
        for (int r = 0; r < reps + 3; r++) {
            cudaMemsetAsync(flush, r, 256u << 20);
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (r >= 3) { sum += ms; best = ms < best ? ms : best; }
        }
        printf("%-44s mean %.2f us  min %.2f us\n", name, sum / reps * 1e3, best * 1e3);
    };
    run("empty 8100 x 256, 4/SM", [&] { k_empty<4><<<n, 256>>>(f, img, 0, range); });
    run("empty 8100 x 256, 8/SM", [&] { k_empty<8><<<n, 256>>>(f, img, 0, range); });
    run("flag read + exit, 5/SM", [&] { k_empty<5><<<n, 256>>>(f, img, 1, range); });
    run("flag read + exit, 8/SM", [&] { k_empty<8><<<n, 256>>>(f, img, 1, range); });
    run("flag + 192-thread store, 5/SM", [&] { k_empty<5><<<n, 256>>>(f, img, 2, range); });
    run("flag + 192-thread store, 8/SM", [&] { k_empty<8><<<n, 256>>>(f, img, 2, range); });
    run("flag + warp-0 store, 5/SM", [&] { k_empty<5><<<n, 256>>>(f, img, 3, range); });
    run("L1 range check, 2/3 exit, rest flag+store", [&] { k_empty<5><<<n, 256>>>(f, img, 4, range); });
    run("plain fill 24.9 MB, 592 x 256", [&] { k_fill<<<592, 256>>>(img, (size_t)n * 192); });
    run("plain fill 24.9 MB, 1184 x 256", [&] { k_fill<<<1184, 256>>>(img, (size_t)n * 192); });
    run("empty 1 x 32 (launch floor)", [&] { k_empty<4><<<1, 32>>>(f, img, 0, range); });
    return 0;
}
