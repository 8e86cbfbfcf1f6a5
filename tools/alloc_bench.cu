// How expensive is growing the stream-ordered pool vs cudaMalloc for multi-GB blocks?
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(){
  cudaFree(0);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool,0); unsigned long long thr=~0ull; cudaMemPoolSetAttribute(pool,cudaMemPoolAttrReleaseThreshold,&thr);
  for (size_t gb : {1,8,32}) {
    size_t bytes = gb<<30; void* p;
    double t0=now(); cudaMalloc(&p,bytes); cudaDeviceSynchronize(); double t1=now();
    cudaMemsetAsync(p,0,bytes,s); cudaStreamSynchronize(s); double t2=now();
    cudaFree(p); double t3=now();
    printf("cudaMalloc      %2zu GB: alloc %.1f ms, first memset %.1f ms, free %.1f ms\n",gb,(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3);
    t0=now(); cudaMallocAsync(&p,bytes,s); cudaStreamSynchronize(s); t1=now();
    cudaMemsetAsync(p,0,bytes,s); cudaStreamSynchronize(s); t2=now();
    cudaFreeAsync(p,s); cudaStreamSynchronize(s); t3=now();
    printf("cudaMallocAsync %2zu GB: alloc %.1f ms, first memset %.1f ms, free %.1f ms\n",gb,(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3);
    t0=now(); cudaMallocAsync(&p,bytes,s); cudaStreamSynchronize(s); t1=now(); cudaFreeAsync(p,s); cudaStreamSynchronize(s);
    printf("cudaMallocAsync %2zu GB again (pool warm): alloc %.1f ms\n",gb,(t1-t0)*1e3);
  }
  return 0;
}
