#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_set(cudaGraphConditionalHandle h, int v) { if (threadIdx.x == 0) cudaGraphSetConditional(h, v); }
__global__ void k_body(int* x) { *x += 1; }
extern "C" int run(int v)
{
    cudaStream_t s; cudaStreamCreate(&s);
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
    k_set<<<1, 32, 0, s>>>(h, v);
    cudaStreamCaptureStatus st; const cudaGraphNode_t* deps; size_t nd; cudaGraph_t cg; unsigned long long id;
    cudaStreamGetCaptureInfo_v2(s, &st, &id, &cg, &deps, &nd);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, cg, deps, nd, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    cudaStream_t s2; cudaStreamCreate(&s2);
    e = cudaStreamBeginCaptureToGraph(s2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin body: %s\n", cudaGetErrorString(e));
    k_body<<<1, 1, 0, s2>>>(d);
    e = cudaStreamEndCapture(s2, nullptr);
    printf("end body: %s\n", cudaGetErrorString(e));
    e = cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
    printf("update deps: %s\n", cudaGetErrorString(e));
    k_body<<<1, 1, 0, s>>>(d);
    e = cudaStreamEndCapture(s, &g);
    printf("end: %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, cudaGraphInstantiateFlagUseNodePriority);
    printf("inst: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int hv; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost);
    printf("v=%d result=%d (expect %d)\n", v, hv, v ? 2 : 1);
    return hv;
}
int main() { run(0); run(1); return 0; }
