#include <cstdio>
#include <cstdint>
__global__ void k(const uint32_t* src, uint32_t* out, int hint) {
    __shared__ uint2 buf[64];
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    uint32_t sa = (uint32_t) __cvta_generic_to_shared(&buf[threadIdx.x].x);
    if (hint) asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" :: "r"(sa), "l"(src + threadIdx.x * 3), "l"(pol) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(src + threadIdx.x * 3) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    out[threadIdx.x] = buf[threadIdx.x].x;
}
int main() {
    uint32_t *s, *o; cudaMalloc(&s, 4096); cudaMalloc(&o, 256);
    uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = i * 7; cudaMemcpy(s, h, 4096, cudaMemcpyHostToDevice);
    for (int hint = 0; hint < 2; hint++) {
        k<<<1, 32>>>(s, o, hint);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t r[32]; cudaMemcpy(r, o, 128, cudaMemcpyDeviceToHost);
        printf("hint=%d: %s r[5]=%u (expect %u)\n", hint, cudaGetErrorString(e), r[5], 15 * 7);
        if (e != cudaSuccess) break;
    }
}
