// Probe: which 2-D TMA box shapes / element types work (developer experiment, not product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int bytes, int c0, int c1, unsigned char *out, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ((bytes + 127) / 128 * 128));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (mode == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(smem)), "l"(&tmap), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char **argv) {
    int esz = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), W = atoi(argv[4]), H = atoi(argv[5]), c0 = atoi(argv[6]), c1 = atoi(argv[7]);
    int mode = argc > 8 ? atoi(argv[8]) : 0;
    size_t pitch = ((size_t)W * esz + 511) / 512 * 512;
    unsigned char *src, *out;
    cudaMalloc(&src, pitch * H); cudaMemset(src, 1, pitch * H);
    int bytes = bw * bh * esz;
    cudaMalloc(&out, bytes);
    void *fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    cuuint64_t gd[2] = {(cuuint64_t)W, (cuuint64_t)H}, gs[1] = {pitch};
    cuuint32_t bx[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CUtensorMapDataType dt = esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUresult r = ((Enc)fn)(&tm, dt, 2, src, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    size_t smem = (bytes + 127) / 128 * 128 + 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 256, smem>>>(tm, bytes, c0, c1, out, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("esz=%d box=%dx%d (%d B, inner %d B) tensor=%dx%d at (%d,%d) mode=%d: encode=%d run=%s\n", esz, bw, bh, bytes, bw * esz, W, H, c0, c1, mode, (int)r, cudaGetErrorName(e));
    return 0;
}
