#include <cstdio>
#include <cuda_runtime.h>
__global__ void body(int *ctr, cudaGraphConditionalHandle h) { int v = atomicAdd(ctr, 1); if (threadIdx.x == 0) cudaGraphSetConditional(h, v + 1 < 5 ? 1u : 0u); }
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  int *ctr; cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p); printf("add %s\n", cudaGetErrorString(e));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal); printf("begin %s\n", cudaGetErrorString(e));
  body<<<1, 1, 0, st>>>(ctr, h);
  e = cudaStreamEndCapture(st, nullptr); printf("end %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %s\n", cudaGetErrorString(e));
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int hc; cudaMemcpy(&hc, ctr, 4, cudaMemcpyDeviceToHost); printf("iterations %d (%s)\n", hc, cudaGetErrorString(cudaGetLastError()));
}
