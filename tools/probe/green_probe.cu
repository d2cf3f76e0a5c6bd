// probe: do green contexts (SM partitions) work here, with runtime launches, shared memory and cross-stream events?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <set>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s; cuGetErrorString(r_, &s); printf("%s -> %s\n", #x, s); return 1; } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)
__global__ void who(int *out, int spin) {
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long t0 = clock64(); while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0) out[blockIdx.x] = (int)sm;
}
int main() {
    RK(cudaSetDevice(0)); RK(cudaFree(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUdevResource sm; CK(cuDeviceGetDevResource(dev, &sm, CU_DEV_RESOURCE_TYPE_SM));
    printf("SMs %u\n", sm.sm.smCount);
    for (unsigned want : {48u, 64u, 72u}) {
        CUdevResource grp, rem; unsigned nb = 1;
        CK(cuDevSmResourceSplitByCount(&grp, &nb, &sm, &rem, 0, want));
        printf("split want %u -> group %u remaining %u (nb %u)\n", want, grp.sm.smCount, rem.sm.smCount, nb);
        CUdevResourceDesc d1, d2; CK(cuDevResourceGenerateDesc(&d1, &grp, 1)); CK(cuDevResourceGenerateDesc(&d2, &rem, 1));
        CUgreenCtx g1, g2; CK(cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM)); CK(cuGreenCtxCreate(&g2, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM));
        CUstream s1, s2; CK(cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0)); CK(cuGreenCtxStreamCreate(&s2, g2, CU_STREAM_NON_BLOCKING, 0));
        int *o1, *o2; RK(cudaMalloc(&o1, 4096 * 4)); RK(cudaMalloc(&o2, 4096 * 4));
        cudaEvent_t e; RK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaEvent_t t0, t1; RK(cudaEventCreate(&t0)); RK(cudaEventCreate(&t1));
        cudaStream_t main_s; RK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
        RK(cudaEventRecord(t0, main_s));
        RK(cudaEventRecord(e, main_s));
        RK(cudaStreamWaitEvent((cudaStream_t)s1, e, 0)); RK(cudaStreamWaitEvent((cudaStream_t)s2, e, 0));
        who<<<2048, 128, 0, (cudaStream_t)s1>>>(o1, 20000); RK(cudaGetLastError());
        who<<<2048, 128, 0, (cudaStream_t)s2>>>(o2, 20000); RK(cudaGetLastError());
        cudaEvent_t d1e, d2e; RK(cudaEventCreateWithFlags(&d1e, cudaEventDisableTiming)); RK(cudaEventCreateWithFlags(&d2e, cudaEventDisableTiming));
        RK(cudaEventRecord(d1e, (cudaStream_t)s1)); RK(cudaEventRecord(d2e, (cudaStream_t)s2));
        RK(cudaStreamWaitEvent(main_s, d1e, 0)); RK(cudaStreamWaitEvent(main_s, d2e, 0));
        RK(cudaEventRecord(t1, main_s));
        RK(cudaStreamSynchronize(main_s));
        float ms; RK(cudaEventElapsedTime(&ms, t0, t1));
        int h1[2048], h2[2048]; RK(cudaMemcpy(h1, o1, sizeof h1, cudaMemcpyDeviceToHost)); RK(cudaMemcpy(h2, o2, sizeof h2, cudaMemcpyDeviceToHost));
        std::set<int> a(h1, h1 + 2048), b(h2, h2 + 2048); int overlap = 0; for (int x : a) overlap += b.count(x);
        printf("  kernel on group ran on %zu SMs, on remaining %zu SMs, overlap %d, both %.3f ms\n", a.size(), b.size(), overlap, ms);
        CK(cuStreamDestroy(s1)); CK(cuStreamDestroy(s2)); CK(cuGreenCtxDestroy(g1)); CK(cuGreenCtxDestroy(g2));
        cudaFree(o1); cudaFree(o2);
    }
    printf("ok\n");
    return 0;
}
