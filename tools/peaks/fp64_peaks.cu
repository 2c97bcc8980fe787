// FP64 roofline denominators for B200 (sm_100a): DFMA issue peak, DMMA.8x8x4 issue peak,
// shared-memory-fed DMMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
// Prints one JSON object on stdout.  Timing: CUDA events on the launching stream, 3 warm-ups, best of 10.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double x, double y){
  double acc[ILP];
#pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-9+i;
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i],x,y);
  }
  double s=0;
#pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters){
  double a=threadIdx.x*1e-3, b=threadIdx.x*2e-3;
  double c[NACC][2];
#pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
#pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float best_ms(F f){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for(int i=0;i<3;i++) f();
  CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int i=0;i<10;i++){
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best) best=ms;
  }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*256));
  const int iters=20000;
  printf("{\"gpu\":\"%s\",\"sms\":%d", p.name, sms);
  {
    int blocks=sms*4;
    float ms=best_ms([&]{ dfma_kernel<16><<<blocks,256>>>(out,iters,1.0000001,1e-9); });
    double fl=2.0*16*iters*256.0*blocks;
    printf(",\"dfma_tflops\":%.3f", fl/ms*1e-9);
    // sustained: ~2 s back to back
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int reps=(int)(2000.0f/ms)+1; CK(cudaEventRecord(e0));
    for(int i=0;i<reps;i++) dfma_kernel<16><<<blocks,256>>>(out,iters,1.0000001,1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t,e0,e1));
    printf(",\"dfma_tflops_sustained\":%.3f", fl*reps/t*1e-9);
  }
  {
    int blocks=sms*4;
    float ms=best_ms([&]{ dmma_kernel<8><<<blocks,256>>>(out,iters); });
    double fl=2.0*8*8*4*8.0*iters*(256/32)*(double)blocks;
    printf(",\"dmma884_tflops\":%.3f", fl/ms*1e-9);
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int reps=(int)(2000.0f/ms)+1; CK(cudaEventRecord(e0));
    for(int i=0;i<reps;i++) dmma_kernel<8><<<blocks,256>>>(out,iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t,e0,e1));
    printf(",\"dmma884_tflops_sustained\":%.3f", fl*reps/t*1e-9);
  }
  {
    int blocks=sms*2;
    float ms=best_ms([&]{ dmma_kernel<8><<<blocks,128>>>(out,iters); });
    double fl=2.0*8*8*4*8.0*iters*(128/32)*(double)blocks;
    printf(",\"dmma884_tflops_8warps_per_sm\":%.3f", fl/ms*1e-9);
  }
  {
    int blocks=sms;
    float ms=best_ms([&]{ dmma_kernel<8><<<blocks,128>>>(out,iters); });
    double fl=2.0*8*8*4*8.0*iters*(128/32)*(double)blocks;
    printf(",\"dmma884_tflops_4warps_per_sm\":%.3f", fl/ms*1e-9);
  }
  printf("}\n");
  return 0;
}
