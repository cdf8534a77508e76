#include <cuda_runtime.h>
#include <stdio.h>
__global__ void body(cudaGraphConditionalHandle h, int *counter, float *x) {
    x[threadIdx.x] += 1.f;
    if (threadIdx.x == 0) { int c = *counter - 1; *counter = c; cudaGraphSetConditional(h, c > 0); }
}
__global__ void head(cudaGraphConditionalHandle h, const int *counter) { cudaGraphSetConditional(h, *counter > 0); }
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int *cnt; float *x; cudaMalloc(&cnt, 4); cudaMalloc(&x, 128);
    cudaMemset(x, 0, 128);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 0, 0);
    cudaGraphNode_t nhead, nwhile;
    cudaKernelNodeParams kp = {}; void *args[] = {&h, &cnt}; kp.func = (void*)head; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    printf("%d\n", cudaGraphAddKernelNode(&nhead, g, nullptr, 0, &kp));
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    printf("%d\n", cudaGraphAddNode(&nwhile, g, &nhead, 1, &p));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    printf("%d\n", cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    body<<<1, 32, 0, st>>>(h, cnt, x);
    cudaGraph_t tmp; printf("%d\n", cudaStreamEndCapture(st, &tmp));
    cudaGraphExec_t ex; printf("%d\n", cudaGraphInstantiate(&ex, g, 0));
    int n = 7; cudaMemcpy(cnt, &n, 4, cudaMemcpyHostToDevice);
    cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    float hx; cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost); printf("x=%f (expect 7) err=%d\n", hx, cudaGetLastError());
}
