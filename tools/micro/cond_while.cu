#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *c, cudaGraphConditionalHandle h) { int v = ++(*c); cudaGraphSetConditional(h, v < 5 ? 1u : 0u); }
int main(){
  cudaStream_t s, s2; cudaStreamCreate(&s); cudaStreamCreate(&s2);
  int *c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; 
  printf("%d\n", (int)cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t n; printf("%d\n", (int)cudaGraphAddNode(&n, g, nullptr, 0, &p));
  cudaGraph_t b = p.conditional.phGraph_out[0];
  cudaStreamBeginCaptureToGraph(s2, b, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  body<<<1,1,0,s2>>>(c, h);
  cudaStreamEndCapture(s2, nullptr);
  cudaGraphExec_t e; printf("%d\n", (int)cudaGraphInstantiate(&e, g, 0));
  cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
  int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost); printf("c=%d\n", hc);
}
