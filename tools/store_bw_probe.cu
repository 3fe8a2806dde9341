// store_bw_probe.cu -- how fast can ONE SM (and 148 of them) write zeros to HBM?
//   mode 0: W warps per CTA issue st.global.cs.v4 (512 B per warp instruction)
//   mode 1: T threads per CTA issue cp.async.bulk shared->global copies of CHUNK bytes from one zero buffer
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_bw_probe tools/store_bw_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__global__ void st_kernel(float4* out, size_t per_cta_f4) {
    float4* base = out + (size_t)blockIdx.x * per_cta_f4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = threadIdx.x; i < per_cta_f4; i += blockDim.x) __stcs(base + i, z);
}

__global__ void bulk_kernel(uint8_t* out, size_t per_cta_bytes, uint32_t chunk, int issuers) {
    extern __shared__ __align__(128) uint8_t zero[];
    for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<float4*>(zero + i) = make_float4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < issuers) {
        const int me = threadIdx.x >> 5;
        uint8_t* base = out + (size_t)blockIdx.x * per_cta_bytes;
        uint32_t saddr = (uint32_t)__cvta_generic_to_shared(zero);
        for (size_t off = (size_t)me * chunk; off < per_cta_bytes; off += (size_t)issuers * chunk)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(saddr), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    const int ctas = 148;
    const size_t per_cta = 8u << 20;   // 8 MiB per CTA -> 1.16 GiB total, far beyond L2
    uint8_t* buf;
    cudaMalloc(&buf, per_cta * ctas);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("mode,param,ctas,GB/s total,GB/s per SM\n");
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            st_kernel<<<ctas, warps * 32>>>((float4*)buf, per_cta / 16);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        double gbs = per_cta * ctas / (time_ms(e0, e1) * 1e-3) / 1e9;
        printf("st.v4,warps=%d,%d,%.0f,%.1f\n", warps, ctas, gbs, gbs / ctas);
    }
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (uint32_t chunk : {1024u, 4096u, 16384u, 65536u}) {
        for (int issuers : {1, 4}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                bulk_kernel<<<ctas, 128, chunk>>>(buf, per_cta, chunk, issuers);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            double gbs = per_cta * ctas / (time_ms(e0, e1) * 1e-3) / 1e9;
            printf("bulk,chunk=%u issuers=%d,%d,%.0f,%.1f\n", chunk, issuers, ctas, gbs, gbs / ctas);
        }
    }
    // one SM alone: is the limit per SM or chip-wide?
    for (int warps : {1, 4, 16}) {
        cudaEventRecord(e0);
        st_kernel<<<1, warps * 32>>>((float4*)buf, per_cta / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        double gbs = per_cta / (time_ms(e0, e1) * 1e-3) / 1e9;
        printf("st.v4 single CTA,warps=%d,1,%.1f,%.1f\n", warps, gbs, gbs);
    }
    cudaEventRecord(e0);
    bulk_kernel<<<1, 128, 16384>>>(buf, per_cta, 16384, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    printf("bulk single CTA,chunk=16384 issuers=1,1,%.1f,-\n", per_cta / (time_ms(e0, e1) * 1e-3) / 1e9);
    cudaError_t err = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
