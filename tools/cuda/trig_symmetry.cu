// Developer check (run on the B200): are libdevice's cosf / sinf exactly even / odd, and does sincosf return the very
// bits of separate sinf / cosf calls, for ALL 2^32 float inputs?  If so, device_trig() may evaluate one sincosf instead of
// the four calls the reference issues.   nvcc -arch=sm_100a -o trig_symmetry trig_symmetry.cu && ./trig_symmetry
#include <cstdio>
#include <cstdint>
__global__ void check(unsigned long long* bad) {
    const uint64_t n = 1ull << 32;
    unsigned long long b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        const float c = cosf(x), s = sinf(x), cn = cosf(-x), sn = sinf(-x);
        float ss, cc;
        sincosf(x, &ss, &cc);
        const bool nan_c = c != c, nan_s = s != s;
        if (nan_c ? !(cn != cn) : (__float_as_uint(cn) != __float_as_uint(c))) ++b0;
        if (nan_s ? !(sn != sn) : (__float_as_uint(sn) != __float_as_uint(-s))) ++b1;
        if (nan_c ? !(cc != cc) : (__float_as_uint(cc) != __float_as_uint(c))) ++b2;
        if (nan_s ? !(ss != ss) : (__float_as_uint(ss) != __float_as_uint(s))) ++b3;
    }
    atomicAdd(&bad[0], b0); atomicAdd(&bad[1], b1); atomicAdd(&bad[2], b2); atomicAdd(&bad[3], b3);
}
int main() {
    unsigned long long* d; unsigned long long h[4] = {0, 0, 0, 0};
    cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    check<<<148 * 8, 256>>>(d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mismatches over 2^32 inputs: cosf(-x) vs cosf(x): %llu   sinf(-x) vs -sinf(x): %llu   sincosf.c vs cosf: %llu   sincosf.s vs sinf: %llu\n", h[0], h[1], h[2], h[3]);
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
