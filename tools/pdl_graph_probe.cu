// pdl_graph_probe.cu -- does stream capture accept a programmatic-launch kernel whose dependency set holds BOTH the
// previous kernel of its own stream and a kernel of a forked stream (joined by an event)?  Prints the captured graph's
// edges with their types and the replay time of a 3-kernel chain + side kernel, next to the same chain launched eagerly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/pdl_graph_probe tools/pdl_graph_probe.cu && tools/pdl_graph_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void work(float* p, int n, int spin) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = p[i];
        for (int k = 0; k < spin; ++k) v = v * 1.0001f + 0.5f;
        p[i] = v;
    }
}

static cudaError_t launch(cudaStream_t st, bool pdl, float* p, int n, int spin) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, work, p, n, spin);
}

int main() {
    const int n = 1 << 20;
    float *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t fork[8], join[8];
    for (int i = 0; i < 8; ++i) { CK(cudaEventCreateWithFlags(&fork[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming)); }
    const int steps = 6;
    auto enqueue = [&](bool side_pdl) -> int {
        for (int i = 0; i < steps; ++i) {
            if (i >= 2) CK(cudaStreamWaitEvent(s1, join[(i - 2) % 8], 0));      // join of the side kernel of step i-2
            CK(launch(s1, true, a, n, 20));       // "prepare"
            CK(launch(s1, true, a, n, 200));      // "forward"
            CK(launch(s1, true, a, n, 50));       // "backward"
            CK(cudaEventRecord(fork[i % 8], s1));
            CK(cudaStreamWaitEvent(s2, fork[i % 8], 0));
            CK(launch(s2, side_pdl, b, n, 100));  // "exchange" on the side stream
            CK(cudaEventRecord(join[i % 8], s2));
        }
        CK(cudaStreamWaitEvent(s1, join[(steps - 1) % 8], 0));
        CK(cudaStreamWaitEvent(s1, join[(steps - 2) % 8], 0));
        return 0;
    };
    for (int side_pdl = 0; side_pdl < 2; ++side_pdl) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal));
        if (enqueue(side_pdl != 0)) { printf("capture enqueue failed (side_pdl=%d)\n", side_pdl); cudaStreamEndCapture(s1, &g); continue; }
        cudaError_t e = cudaStreamEndCapture(s1, &g);
        if (e != cudaSuccess) { printf("end capture failed (side_pdl=%d): %s\n", side_pdl, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        size_t ne = 0;
        CK(cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne));
        std::vector<cudaGraphNode_t> from(ne), to(ne);
        std::vector<cudaGraphEdgeData> ed(ne);
        CK(cudaGraphGetEdges_v2(g, from.data(), to.data(), ed.data(), &ne));
        int prog = 0, full = 0;
        for (size_t i = 0; i < ne; ++i) (ed[i].type == cudaGraphDependencyTypeProgrammatic ? prog : full)++;
        printf("side_pdl=%d: captured, %zu edges: %d programmatic, %d full\n", side_pdl, ne, prog, full);
        e = cudaGraphInstantiate(&ge, g, 0);
        if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); continue; }
        cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s1));
        CK(cudaEventRecord(t0, s1));
        for (int r = 0; r < 20; ++r) CK(cudaGraphLaunch(ge, s1));
        CK(cudaEventRecord(t1, s1)); CK(cudaStreamSynchronize(s1));
        float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("side_pdl=%d: graph replay %.2f us per step\n", side_pdl, 1e3 * ms / (20 * steps));
    }
    {   // eager, same structure
        cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        for (int w = 0; w < 3; ++w) if (enqueue(false)) return 1;
        CK(cudaEventRecord(t0, s1));
        for (int r = 0; r < 20; ++r) if (enqueue(false)) return 1;
        CK(cudaEventRecord(t1, s1)); CK(cudaStreamSynchronize(s1));
        float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("eager fork/join: %.2f us per step\n", 1e3 * ms / (20 * steps));
        // eager, one stream, no side kernel
        CK(cudaEventRecord(t0, s1));
        for (int r = 0; r < 20 * steps; ++r) { CK(launch(s1, true, a, n, 20)); CK(launch(s1, true, a, n, 200)); CK(launch(s1, true, a, n, 50)); }
        CK(cudaEventRecord(t1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("eager single stream, 3 kernels (no side kernel): %.2f us per step\n", 1e3 * ms / (20 * steps));
        CK(cudaEventRecord(t0, s1));
        for (int r = 0; r < 20 * steps; ++r) { CK(launch(s1, true, a, n, 20)); CK(launch(s1, true, a, n, 200)); CK(launch(s1, true, a, n, 50)); CK(launch(s1, true, b, n, 100)); }
        CK(cudaEventRecord(t1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("eager single stream, 4 kernels in line: %.2f us per step\n", 1e3 * ms / (20 * steps));
    }
    printf("done\n");
    return 0;
}
