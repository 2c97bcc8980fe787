#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int* counter) { atomicAdd(counter, 1); }
__global__ void k_check(cudaGraphConditionalHandle h, int* counter, int maxit) {
  cudaGraphSetConditional(h, *counter < maxit ? 1 : 0);
}
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
  printf("addnode %s\n", cudaGetErrorString(e));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  printf("begin %s\n", cudaGetErrorString(e));
  k_body<<<1, 1, 0, s>>>(d);
  k_check<<<1, 1, 0, s>>>(h, d, 7);
  cudaGraph_t dummy; e = cudaStreamEndCapture(s, &dummy);
  printf("end %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
  printf("inst %s\n", cudaGetErrorString(e));
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  int hc; cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost); printf("count %d\n", hc);
  cudaMemset(d, 0, 4);
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost); printf("count again %d %s\n", hc, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
