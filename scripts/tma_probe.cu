// Debug probe for the TMA path of diffusion.cu: loads one box with cp.async.bulk.tensor.2d and checks it.
// usage: tma_probe <f64|f32> <boxJ> <boxI> <gx> <gy> <c0> <c1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap map, unsigned char *out, int bytes, int c0, int c1) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~uintptr_t(127));
    uint64_t *bar = (uint64_t *)(base + ((bytes + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(base)), "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
        ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = base[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    if (argc < 8) return 2;
    bool f64 = !strcmp(argv[1], "f64");
    int boxJ = atoi(argv[2]), boxI = atoi(argv[3]);
    long gx = atol(argv[4]), gy = atol(argv[5]);
    int c0 = atoi(argv[6]), c1 = atoi(argv[7]);
    int l2p = argc > 8 ? atoi(argv[8]) : 2, oob = argc > 9 ? atoi(argv[9]) : 0;
    int es = f64 ? 8 : 4;
    size_t n = (size_t)gx * gy;
    void *d;
    cudaMalloc(&d, n * es);
    double *h64 = (double *)malloc(n * 8);
    float *h32 = (float *)h64;
    for (size_t i = 0; i < n; ++i) { if (f64) h64[i] = (double)i; else h32[i] = (float)i; }
    cudaMemcpy(d, h64, n * es, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)gy, (cuuint64_t)gx};
    cuuint64_t gstride[1] = {(cuuint64_t)gy * es};
    cuuint32_t box[2] = {(cuuint32_t)boxJ, (cuuint32_t)boxI};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d,
                                     gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     (CUtensorMapL2promotion)l2p, (CUtensorMapFloatOOBfill)oob);
    printf("encode rc=%d\n", (int)r);
    if (r) return 1;
    int bytes = boxJ * boxI * es;
    unsigned char *out;
    cudaMalloc(&out, bytes);
    size_t smem = ((bytes + 127) / 128) * 128 + 128 + 8;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(map, out, bytes, c0, c1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    void *ho = malloc(bytes);
    cudaMemcpy(ho, out, bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < boxI; ++i)
        for (int j = 0; j < boxJ; ++j) {
            long gi = c1 + i, gj = c0 + j;
            double want = (gi >= 0 && gi < gx && gj >= 0 && gj < gy) ? (double)(gi * gy + gj) : 0.0;
            double got = f64 ? ((double *)ho)[i * boxJ + j] : (double)((float *)ho)[i * boxJ + j];
            if (got != want) ++bad;
        }
    printf("mismatches=%ld of %d\n", bad, boxI * boxJ);
    return bad != 0;
}
