// Developer microbenchmark (B200): HBM write bandwidth of the zero-fill patterns considered for iou_tile_kernel.
//   A  592 resident CTAs, each bulk-stores its own contiguous 150 KB region (what the kernel does), many waves
//   B  same bytes, but the 4 KB chunks are interleaved over the CTAs (chunk k -> CTA k % grid): a narrow moving window
//   C  plain 16-byte stores, grid-stride (torch.zero_ style)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_patterns fill_patterns.cu && ./fill_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ZB = 4096;
__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(__cvta_generic_to_global(g)), "r"((unsigned)__cvta_generic_to_shared(s)), "r"(n) : "memory");
}
__global__ void __launch_bounds__(256, 4) fill_own(char* out, size_t tile_bytes, int tiles) {
    __shared__ __align__(128) float4 zero[ZB / 16];
    extern __shared__ char pad[];   // occupy smem like the real kernel (4 CTAs/SM)
    zero[threadIdx.x] = make_float4(0, 0, 0, 0);
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        if (threadIdx.x < 32) {
            char* dst = out + (size_t)t * tile_bytes;
            for (size_t off = (size_t)threadIdx.x * ZB; off < tile_bytes; off += 32 * ZB) bulk_store(dst + off, zero, (unsigned)min((size_t)ZB, tile_bytes - off));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256, 4) fill_interleaved(char* out, size_t total) {
    __shared__ __align__(128) float4 zero[ZB / 16];
    extern __shared__ char pad[];
    zero[threadIdx.x] = make_float4(0, 0, 0, 0);
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        const size_t chunks = total / ZB;
        // lane l of CTA b takes chunks (k * gridDim.x + b) * 32 + l  -> at any time the grid writes a window of grid*32 chunks
        for (size_t k = 0;; ++k) {
            const size_t c = (k * gridDim.x + blockIdx.x) * 32 + threadIdx.x;
            if (c - threadIdx.x >= chunks) break;
            if (c < chunks) bulk_store(out + c * ZB, zero, ZB);
            if ((k & 7) == 7) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void fill_stg(float4* out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_float4(0, 0, 0, 0);
}
int main() {
    const size_t tile = 384 * 100 * 4, tiles = 8800, total = tile * tiles;   // 1.35 GB
    char* d; cudaMalloc(&d, total);
    cudaFuncSetAttribute(fill_own, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaFuncSetAttribute(fill_interleaved, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); for (int i = 0; i < 5; ++i) fill_own<<<592, 256, 48 * 1024>>>(d, tile, (int)tiles); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("A own contiguous tile per CTA (bulk):   %.1f us  %.2f TB/s\n", ms / 5 * 1e3, total / (ms / 5 * 1e-3) / 1e12);
        cudaEventRecord(a); for (int i = 0; i < 5; ++i) fill_interleaved<<<592, 256, 48 * 1024>>>(d, total); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("B interleaved 4 KB chunks (bulk):       %.1f us  %.2f TB/s\n", ms / 5 * 1e3, total / (ms / 5 * 1e-3) / 1e12);
        cudaEventRecord(a); for (int i = 0; i < 5; ++i) fill_stg<<<148 * 8, 256>>>((float4*)d, total / 16); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("C grid-stride 16-byte stores:           %.1f us  %.2f TB/s\n", ms / 5 * 1e3, total / (ms / 5 * 1e-3) / 1e12);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
