// Separable kernel for axis-aligned reduced angles (0/90/180/270 degrees): the main loop of
// AreaAverageInterpolation::areaAverageInterpolation (Source.cpp:411-579) when every footprint is an axis-aligned
// square, so that the overlap area factorises, A_ij = wx(i) * wy(j), and
//     dst = sum_j wy(j) sum_i wx(i) src[j][i] / (sum wx * sum wy)        (sums over in-image cells, Source.cpp:577).
//
// HBM-bound by contract (each source byte read once, each canvas byte written once).  One CTA = one canvas tile
// TW x TH.  Its source window is staged into shared memory by ONE 2-D TMA tile load (cp.async.bulk.tensor.2d through
// a CUtensorMap, completion on an mbarrier; out-of-image parts of the box are zero-filled by the hardware and carry
// weight 0).  The two banded 1-D passes run out of that tile with the intermediate in registers: for every canvas pixel
// the horizontal taps of each of its source rows (weights in registers, taps read as aligned vectors so that the
// stride-L accesses of a warp do not collide on shared-memory banks), then the vertical taps (weights broadcast from
// shared memory), then the coalesced store.  (A first version wrote the horizontal pass to a shared-memory tile and
// synchronised: 3x the instructions, issue-bound at 64 %; see profiles/.)  Several CTAs are resident per SM, so one
// CTA's TMA load overlaps the other CTAs' arithmetic.
//
// A batch of equally strided images is ONE launch: the tensor map is rank 3 (x, y, image), blockIdx.y = image.
//
// Fast path preconditions (checked by the host launcher, otherwise the direct-tap kernel in aai_kernels.cu runs):
// scale 1, quadrant 0, one channel, 16-byte aligned rows, footprint side small enough for a <=256-wide TMA box.
// TMA tile coordinates must be 16-byte aligned in the innermost dimension: the window origin is rounded down.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "aai_device.cuh"

using namespace aai_dev;

namespace {

constexpr int SEP_THREADS = 256;

struct SepParams {
    int tw, th;  // canvas tile
    int bw, bh;  // TMA box = source window of one tile (elements, rows)
    int tiles_x;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// (x, y, image) box of a rank-3 tensor map: a batch of equally strided images is one tensor
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <typename TA>
__device__ __forceinline__ TA sat01(TA x) {
    return x < (TA)0 ? (TA)0 : (x > (TA)1 ? (TA)1 : x);
}
template <>
__device__ __forceinline__ float sat01<float>(float x) {
    return __saturatef(x);
}

// taps of one canvas coordinate: cells [first, first+MAXT) around the footprint interval [c-h, c+h], weight of tap t =
// |[first+t-1/2, first+t+1/2] ∩ [c-h, c+h]| for in-image cells, 0 otherwise
template <typename TA, int MAXT>
__device__ __forceinline__ void axis_taps(double c, double h, int limit, int &first, TA (&w)[MAXT], TA &sum) {
    first = __double2int_rd(c - h - 0.5) + 1;
    const double e = (double)first - 0.5;
    const TA lo = (TA)((c - h) - e), hi = (TA)((c + h) - e);
    sum = (TA)0;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        TA v = sat01<TA>(hi - (TA)t) - sat01<TA>(lo - (TA)t);
        const int i = first + t;
        v = (i >= 0 && i < limit) ? v : (TA)0;
        w[t] = v;
        sum += v;
    }
}

// aligned vector of source elements in shared memory: 8 bytes for 4-byte elements, otherwise one element
template <typename TI>
struct TapVec {
    static constexpr int N = 1;
    static __device__ __forceinline__ void load(const TI *p, TI (&v)[1]) { v[0] = *p; }
};
template <>
struct TapVec<float> {
    static constexpr int N = 2;
    static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        v[0] = t.x;
        v[1] = t.y;
    }
};

template <typename TI, typename TO, typename TA, int TW, int MAXT>
__global__ void __launch_bounds__(SEP_THREADS)
    separable_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ AaiKernelParams kp,
                         const SepParams sp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int RG = SEP_THREADS / TW;  // row groups
    constexpr int VEC = TapVec<TI>::N;
    constexpr int NV = (MAXT + 2 * (VEC - 1)) / VEC;  // aligned vectors that cover MAXT taps from any start offset
    TI *tile = reinterpret_cast<TI *>(smem_raw);
    const size_t tile_bytes = ((size_t)sp.bw * sp.bh * sizeof(TI) + 127) / 128 * 128;
    TA *wyw = reinterpret_cast<TA *>(smem_raw + tile_bytes);                         // [th][MAXT]
    TA *wys = wyw + (size_t)sp.th * MAXT;                                            // [th] 1/sum
    int *wyf = reinterpret_cast<int *>(wys + sp.th);                                 // [th] first row (tile-relative)
    uint64_t *bar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(wyf + sp.th) + 7) / 8 * 8);

    const int tid = threadIdx.x;
    const int tx = blockIdx.x % sp.tiles_x, ty = blockIdx.x / sp.tiles_x;
    const int x0 = tx * TW, y0 = kp.row0 + ty * sp.th;
    const double h = kp.shape.half;

    // source window origin of this tile (expanded frame == source frame on this path)
    double c0x, c0y;
    pixel_centre(kp, x0, y0, c0x, c0y);
    // TMA needs the box to start on a 16-byte boundary of the innermost dimension (measured on B200: a misaligned
    // start coordinate raises cudaErrorIllegalInstruction), so the window origin is rounded down to ALIGN elements
    constexpr int ALIGN = 16 / (int)sizeof(TI);
    const int ox_raw = __double2int_rd(c0x - h - 0.5);
    const int ox = (ox_raw >= 0 ? ox_raw / ALIGN : -((-ox_raw + ALIGN - 1) / ALIGN)) * ALIGN;
    const int oy = __double2int_rd(c0y - h - 0.5);

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(sp.bw * sp.bh * sizeof(TI)));
        tma_load_3d(tile, &tmap, bar, ox, oy - kp.src_y0, (int)blockIdx.y);
    }

    // while the TMA load is in flight: per-column weights (registers) and per-row weights (shared memory)
    const int xo = tid % TW, rg = tid / TW;
    const int x = x0 + xo;
    int cfirst = 0;
    TA wx[MAXT], sumx = (TA)0;
    if (x < kp.dst_w) {
        double cx, cy;
        pixel_centre(kp, x, y0, cx, cy);
        axis_taps<TA, MAXT>(cx, h, kp.mod_w, cfirst, wx, sumx);
        cfirst -= ox;
    } else {
#pragma unroll
        for (int t = 0; t < MAXT; ++t) wx[t] = (TA)0;
    }
    if (tid < sp.th) {
        const int y = y0 + tid;
        TA w[MAXT], s = (TA)0;
        int first = oy;
        if (y < kp.row1) {
            double cx, cy;
            pixel_centre(kp, x0, y, cx, cy);
            axis_taps<TA, MAXT>(cy, h, kp.mod_h, first, w, s);
        } else {
#pragma unroll
            for (int t = 0; t < MAXT; ++t) w[t] = (TA)0;
        }
#pragma unroll
        for (int t = 0; t < MAXT; ++t) wyw[tid * MAXT + t] = w[t];
        wys[tid] = s;
        wyf[tid] = first - oy;
    }
    // clamp the tap window into the box (taps beyond it have weight 0 by construction of the box size), then widen it
    // to whole aligned vectors: weights wv[] = wx[] shifted by the start offset inside the first vector, zero elsewhere
    cfirst = max(0, min(cfirst, sp.bw - MAXT - 2 * (VEC - 1)));
    const int cbase = cfirst / VEC * VEC, shift = cfirst - cbase;
    TA wv[NV * VEC];
#pragma unroll
    for (int q = 0; q < NV * VEC; ++q) {
        TA v = (TA)0;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) v = (q - shift == t) ? wx[t] : v;
        wv[q] = v;
    }
    __syncthreads();  // row weights visible

    mbar_wait(bar, 0);

    if (x < kp.dst_w) {
        for (int yo = rg; yo < sp.th; yo += RG) {
            const int y = y0 + yo;
            if (y >= kp.row1) break;
            const int rf = max(0, min(wyf[yo], sp.bh - MAXT));
            const TI *row = tile + (size_t)rf * sp.bw + cbase;
            TA acc = (TA)0;
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                // horizontal taps of source row rf + t ...
                TA hs = (TA)0;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    TI v[VEC];
                    TapVec<TI>::load(row + q * VEC, v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) hs += wv[q * VEC + e] * (TA)v[e];
                }
                // ... then its vertical tap
                acc += wyw[yo * MAXT + t] * hs;
                row += sp.bw;
            }
            const TA total = sumx * wys[yo];
            const double out = ((double)total > DBL_EPSILON) ? (double)(acc / total) : 0.0;  // Source.cpp:577
            char *drow = (char *)kp.dst + (int64_t)blockIdx.y * kp.dst_batch_stride + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
            store_dst<TO>(drow, x, out);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

template <typename TI>
CUtensorMapDataType tmap_dtype();
template <>
CUtensorMapDataType tmap_dtype<double>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT64; }
template <>
CUtensorMapDataType tmap_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <>
CUtensorMapDataType tmap_dtype<uint8_t>() { return CU_TENSOR_MAP_DATA_TYPE_UINT8; }

// returns cudaErrorNotSupported when the fast path does not apply (the caller then uses the direct-tap kernel)
template <typename TI, typename TO, typename TA, int TW, int MAXT>
cudaError_t launch_sep(const AaiKernelParams &kp, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const double L = 2.0 * kp.shape.half;
    const int esz = (int)sizeof(TI), align = 16 / esz;
    SepParams sp;
    sp.tw = TW;
    sp.th = 32;
    // the window must hold MAXT taps starting at the first cell of the LAST column / row of the tile
    // (+ align-1 columns because the window origin is rounded down to a 16-byte boundary)
    // (+ 2(VEC-1) so that the taps can be read as whole aligned vectors)
    sp.bw = ((int)ceil((TW - 1) * L) + MAXT + 3 + 2 * (TapVec<TI>::N - 1) + (align - 1) + align - 1) / align * align;
    sp.bh = (int)ceil((sp.th - 1) * L) + MAXT + 3;
    if (sp.bw > 256 || sp.bh > 256 || sp.bw < MAXT || sp.bh < MAXT) return cudaErrorNotSupported;
    sp.tiles_x = (kp.dst_w + TW - 1) / TW;
    const int rows = kp.row1 - kp.row0;
    const int tiles_y = (rows + sp.th - 1) / sp.th;
    size_t smem = ((size_t)sp.bw * sp.bh * esz + 127) / 128 * 128;
    smem += ((size_t)sp.th * MAXT + sp.th) * sizeof(TA) + (size_t)sp.th * sizeof(int) + 16;
    if (smem > 200 * 1024) return cudaErrorNotSupported;

    CUtensorMap tmap;
    const int batch = kp.batch > 0 ? kp.batch : 1;
    const cuuint64_t gdim[3] = {(cuuint64_t)kp.src_w, (cuuint64_t)kp.src_rows, (cuuint64_t)batch};
    const cuuint64_t gstride[2] = {(cuuint64_t)kp.src_pitch,
                                   (cuuint64_t)(batch > 1 ? kp.src_batch_stride : kp.src_pitch * (int64_t)kp.src_rows)};
    const cuuint32_t box[3] = {(cuuint32_t)sp.bw, (cuuint32_t)sp.bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (gstride[1] % 16 != 0) return cudaErrorNotSupported;
    const CUresult r = enc(&tmap, tmap_dtype<TI>(), 3, const_cast<void *>(kp.src), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorNotSupported;
    if (getenv("AAI_DEBUG"))
        fprintf(stderr, "[aai] separable TMA: TI=%d bytes TA=%d bytes TW=%d MAXT=%d box %dx%d tiles %dx%d smem %zu src %p pitch %lld "
                "w %d rows %d\n", esz, (int)sizeof(TA), TW, MAXT, sp.bw, sp.bh, sp.tiles_x, tiles_y, smem, kp.src,
                (long long)kp.src_pitch, kp.src_w, kp.src_rows);
    auto kernel = separable_tma_kernel<TI, TO, TA, TW, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<dim3(sp.tiles_x * tiles_y, batch), SEP_THREADS, smem, stream>>>(tmap, kp, sp);
    return cudaGetLastError();
}

template <typename TI, typename TO, typename TA>
cudaError_t launch_sep_shape(const AaiKernelParams &kp, cudaStream_t stream) {
    const double L = 2.0 * kp.shape.half;
    // an interval of length L meets at most ceil(L)+1 unit cells with positive length
    const int taps = (int)ceil(L - 1e-12) + 1;
    cudaError_t e = cudaErrorNotSupported;
    if (taps <= 3) e = launch_sep<TI, TO, TA, 64, 3>(kp, stream);
    if (e == cudaErrorNotSupported && taps <= 4) e = launch_sep<TI, TO, TA, 64, 4>(kp, stream);
    if (e == cudaErrorNotSupported && taps <= 6) e = launch_sep<TI, TO, TA, 64, 6>(kp, stream);
    if (e == cudaErrorNotSupported && taps <= 10) e = launch_sep<TI, TO, TA, 32, 10>(kp, stream);
    return e;
}

template <typename TI, typename TA>
cudaError_t launch_sep_dst(const AaiKernelParams &kp, int dst_dtype, cudaStream_t stream) {
    switch (dst_dtype) {
        case AAI_F64: return launch_sep_shape<TI, double, TA>(kp, stream);
        case AAI_F32: return launch_sep_shape<TI, float, TA>(kp, stream);
        case AAI_U8: return launch_sep_shape<TI, uint8_t, TA>(kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <typename TA>
cudaError_t launch_sep_src(const AaiKernelParams &kp, int src_dtype, int dst_dtype, cudaStream_t stream) {
    switch (src_dtype) {
        case AAI_F64: return launch_sep_dst<double, TA>(kp, dst_dtype, stream);
        case AAI_F32: return launch_sep_dst<float, TA>(kp, dst_dtype, stream);
        case AAI_U8: return launch_sep_dst<uint8_t, TA>(kp, dst_dtype, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// Returns cudaErrorNotSupported (as int) when the TMA fast path does not apply.
int aai_launch_separable_tma(const AaiKernelParams &kp, int arith, int src_dtype, int dst_dtype, void *stream) {
    if (kp.scale != 1 || kp.quadrant != 0 || kp.channels != 1) return (int)cudaErrorNotSupported;
    if ((kp.src_pitch % 16) != 0 || (reinterpret_cast<uintptr_t>(kp.src) % 16) != 0) return (int)cudaErrorNotSupported;
    if (kp.row1 <= kp.row0 || kp.dst_w <= 0) return (int)cudaSuccess;
    cudaStream_t st = (cudaStream_t)stream;
    if (arith == AAI_ARITH_F32 && src_dtype != AAI_F64) return (int)launch_sep_src<float>(kp, src_dtype, dst_dtype, st);
    return (int)launch_sep_src<double>(kp, src_dtype, dst_dtype, st);
}
