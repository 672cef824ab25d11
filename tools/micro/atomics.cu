// Microbenchmark: throughput of global fire-and-forget atomics on B200 (fp64 / fp32 / u64 / u32),
// block-private 8 MB regions, random vs hot addresses.  nvcc -arch=sm_100a -O3 -o atomics atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x){ x+=0x9E3779B97F4A7C15ull; x=(x^(x>>30))*0xBF58476D1CE4E5B9ull; x=(x^(x>>27))*0x94D049BB133111EBull; return x^(x>>31);} 
template<typename T> __global__ void k(T* pool, uint64_t stride, int iters, int mode, uint32_t span){
  T* acc = pool + (uint64_t)blockIdx.x*stride;
  uint64_t s = mix(blockIdx.x*1024+threadIdx.x);
  for(int i=0;i<iters;++i){ s = mix(s);
    uint32_t k;
    if(mode==0) k = (uint32_t)(s % span);                 // uniform random
    else if(mode==1) k = (uint32_t)((s>>20) % 64);        // 64 hot addresses
    else { uint32_t r=(uint32_t)(s>>32); k = (r & (r>>8) & (r>>16)) % span; }  // skewed towards few-bit indices
    atomicAdd(&acc[k], (T)1);
  }
}
template<typename T> void run(const char* name){
  int blocks=148, threads=1024, iters=2000; uint64_t stride=1<<20; T* pool; cudaMalloc(&pool, blocks*stride*sizeof(T)); cudaMemset(pool,0,blocks*stride*sizeof(T));
  for(int mode=0;mode<3;++mode){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<T><<<blocks,threads>>>(pool,stride,100,mode,1<<20); cudaDeviceSynchronize();
    cudaEventRecord(a); k<T><<<blocks,threads>>>(pool,stride,iters,mode,1<<20); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b); double n=(double)blocks*threads*iters; printf("%-6s mode %d: %8.2f ms  %8.2f G atomics/s\n", name, mode, ms, n/ms/1e6); }
  cudaFree(pool);
}
int main(){ run<double>("f64"); run<float>("f32"); run<unsigned long long>("u64"); run<unsigned int>("u32"); return 0; }
