// prints the L2-related device attributes used in DESIGN.md (build: nvcc -o tools/devprops tools/devprops.cu)
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("name=%s sms=%d l2=%d persistingL2CacheMaxSize=%d accessPolicyMaxWindowSize=%d smemPerSM=%zu regsPerSM=%d clock=%d memclk=%d buswidth=%d\n",
           p.name, p.multiProcessorCount, p.l2CacheSize, p.persistingL2CacheMaxSize, p.accessPolicyMaxWindowSize,
           p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor, p.clockRate, p.memoryClockRate, p.memoryBusWidth);
    return 0;
}
