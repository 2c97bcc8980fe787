// Do DFMA (FP64 FMA pipe) and DMMA.8x8x4 (FP64 tensor sub-pipe) issue concurrently on B200?  Half the warps of every CTA run a
// DFMA loop, the other half a DMMA loop; also larger FP64 MMA shapes if ptxas takes them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mixed fp64_mixed.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

// mode 0: all warps DFMA; 1: all DMMA; 2: even warps DFMA, odd warps DMMA; 3: every warp interleaves 16 DFMA with 8 DMMA
__global__ void __launch_bounds__(256) mixed(double* out, int iters, int mode, double x, double y){
  const int warp = threadIdx.x >> 5;
  double acc[16];
  double c[8][2];
  double a=threadIdx.x*1e-3, b=threadIdx.x*2e-3;
#pragma unroll
  for(int i=0;i<16;i++) acc[i]=threadIdx.x*1e-9+i;
#pragma unroll
  for(int i=0;i<8;i++){c[i][0]=0;c[i][1]=0;}
  const bool do_fma = mode==0 || (mode==2 && !(warp&1)) || mode==3;
  const bool do_mma = mode==1 || (mode==2 && (warp&1)) || mode==3;
  for(int it=0; it<iters; ++it){
    if(do_fma){
#pragma unroll
      for(int i=0;i<16;i++) acc[i]=fma(acc[i],x,y);
    }
    if(do_mma){
#pragma unroll
      for(int i=0;i<8;i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<16;i++) s+=acc[i];
#pragma unroll
  for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount, blocks=sms*4;
  double* out; CK(cudaMalloc(&out, sizeof(double)*blocks*256));
  const int iters=20000;
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for(int mode=0; mode<4; ++mode){
    float best=1e30f;
    for(int rep=0; rep<6; ++rep){
      CK(cudaEventRecord(e0)); mixed<<<blocks,256>>>(out,iters,mode,1.0000001,1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(rep>1 && ms<best) best=ms;
    }
    double warps = 8.0*blocks;
    double fma_warps = mode==0?warps: mode==2?warps/2 : mode==3?warps:0;
    double mma_warps = mode==1?warps: mode==2?warps/2 : mode==3?warps:0;
    double fl_fma = 2.0*16*iters*32.0*fma_warps, fl_mma = 2.0*8*8*4*8.0*iters*mma_warps;
    printf("mode %d: %.3f ms  dfma %.2f TF/s  dmma %.2f TF/s  total %.2f TF/s\n", mode, best, fl_fma/best*1e-9, fl_mma/best*1e-9, (fl_fma+fl_mma)/best*1e-9);
  }
  return 0;
}
