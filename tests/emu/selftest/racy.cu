// tests/emu/selftest/racy.cu -- proves that the emulator's ThreadSanitizer mode SEES a missing barrier (and only that).
// Not product code: built and run by tests/emu/selftest.py.
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void k_with_barrier(int *out) {
    __shared__ int x[32];
    if (threadIdx.x < 32) x[threadIdx.x] = (int)threadIdx.x + 1;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[(threadIdx.x + 1) & 31];
}

__global__ void k_without_barrier(int *out) {
    __shared__ int y[32];
    if (threadIdx.x < 32) y[threadIdx.x] = (int)threadIdx.x + 1;
    // (no __syncthreads(): thread t reads what thread t + 1 wrote)
    out[blockIdx.x * blockDim.x + threadIdx.x] = y[(threadIdx.x + 1) & 31];
}

__global__ void k_warp_sync_only(int *out) {
    __shared__ int z[64];
    z[threadIdx.x] = (int)threadIdx.x;
    __syncwarp();  // orders the lanes of ONE warp ...
    out[blockIdx.x * blockDim.x + threadIdx.x] = z[threadIdx.x ^ 1];  // ... fine: the partner is in my warp
}

__global__ void k_global_race(int *out) {
    out[0] = (int)threadIdx.x;  // every thread stores to the same word
}

__global__ void k_inter_cta_race(int *out) {
    if (threadIdx.x == 0) out[0] = (int)blockIdx.x;  // one thread per CTA, the same word: a race BETWEEN CTAs
}

int main(int argc, char **argv) {
    int *d = nullptr;
    cudaMalloc(&d, 128 * sizeof(int));
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) k_with_barrier<<<2, 64>>>(d);
    if (which == 1) k_without_barrier<<<2, 64>>>(d);
    if (which == 2) k_warp_sync_only<<<2, 64>>>(d);
    if (which == 3) k_global_race<<<1, 64>>>(d);
    if (which == 4) k_inter_cta_race<<<2, 32>>>(d);
    cudaDeviceSynchronize();
    int h[64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("done %d %d\n", which, h[0]);
    cudaFree(d);
    return 0;
}
