// CUDA kernels (sm_100a) of the area-average hot path: the main loop of
// AreaAverageInterpolation::areaAverageInterpolation (Source.cpp:411-579).
//
// Formulation (NOT the reference's 16-segment-test classifier):
//   * canvas pixel (x,y) -> footprint centre C in expanded-source coordinates (same FP64 expression as 212-219);
//     the footprint is the square |u|<=h, |v|<=h with u = (p-C).(c,-s), v = (p-C).(s,c), h = L/2.
//   * for every expanded source pixel (unit cell) in the footprint's bounding box, the overlap area is
//         A = 1/2 * sum over the 4 cell sides of  dist(V, side) * |side ∩ footprint|
//     (Green's theorem with the origin at the footprint vertex V nearest to the cell: the two footprint edges
//     through V contribute nothing, the two far edges cannot reach the cell because L > sqrt(2)).
//     |side ∩ footprint| comes from the footprint's chord on the grid line, computed once per grid line.
//     This single closed form reproduces all nine shapes of Source.cpp:1052-1401 in general position.
//   * the reference's shape 2 / shape 4 leg quirk (1055-1062) is then substituted: when the only
//     footprint edge crossing the cell is a left/right edge (direction (s,c)) that passes completely through it
//     and isolates exactly one cell corner at distance d, the triangle is 1/2 (1-d/c)(1-d/s).
//     Decisions are made in FP64 from FP64 geometry.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "aai_cell.cuh"
#include "aai_internal.h"

namespace {

constexpr int TILE_W = 16;
constexpr int TILE_H = 16;

template <typename T>
struct SrcLoad;
template <>
struct SrcLoad<double> {
    static __device__ __forceinline__ double get(const void *row, int idx) { return __ldg((const double *)row + idx); }
};
template <>
struct SrcLoad<float> {
    static __device__ __forceinline__ double get(const void *row, int idx) {
        return (double)__ldg((const float *)row + idx);
    }
};
template <>
struct SrcLoad<uint8_t> {
    static __device__ __forceinline__ double get(const void *row, int idx) {
        return (double)__ldg((const uint8_t *)row + idx);
    }
};

template <typename T>
__device__ __forceinline__ void store_dst(void *row, int idx, double v);
template <>
__device__ __forceinline__ void store_dst<double>(void *row, int idx, double v) {
    ((double *)row)[idx] = v;
}
template <>
__device__ __forceinline__ void store_dst<float>(void *row, int idx, double v) {
    ((float *)row)[idx] = (float)v;
}
template <>
__device__ __forceinline__ void store_dst<uint8_t>(void *row, int idx, double v) {
    // the reference defines no 8-bit store; documented rule: round half up, saturate to [0,255]
    double r = floor(v + 0.5);
    r = fmin(fmax(r, 0.0), 255.0);
    ((uint8_t *)row)[idx] = (uint8_t)(int)r;
}

// expanded + quadrant-rotated pixel (mx,my) -> original source pixel (inverse of Source.cpp:163-168)
__device__ __forceinline__ void mod_to_src(const AaiKernelParams &kp, int mx, int my, int &sx, int &sy) {
    int ex, ey;
    switch (kp.quadrant) {
        case 0: ex = mx; ey = my; break;
        case 1: ex = my; ey = kp.mod_w - 1 - mx; break;
        case 2: ex = kp.mod_w - 1 - mx; ey = kp.mod_h - 1 - my; break;
        default: ex = kp.mod_h - 1 - my; ey = mx; break;
    }
    if (kp.scale == 1) {
        sx = ex;
        sy = ey;
    } else {
        sx = ex / kp.scale;
        sy = ey / kp.scale;
    }
}

// canvas pixel centre, evaluated with the reference's operand order and no FMA contraction (212-219)
__device__ __forceinline__ void pixel_centre(const AaiKernelParams &kp, int x, int y, double &cx, double &cy) {
    const double u = __dadd_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)x, kp.off_ix), kp.side), kp.iso_x), kp.off_x);
    const double v = __dadd_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)y, kp.off_iy), kp.side), kp.iso_y), kp.off_y);
    cx = __dadd_rn(__dadd_rn(__dmul_rn(u, kp.shape.cs), __dmul_rn(v, kp.shape.sn)), kp.iso_x);
    cy = __dadd_rn(__dadd_rn(__dmul_rn(-u, kp.shape.sn), __dmul_rn(v, kp.shape.cs)), kp.iso_y);
}

// the reference's clamped search window (426-429)
__device__ __forceinline__ void search_window(const AaiKernelParams &kp, double cx, double cy, int &x0, int &x1,
                                              int &y0, int &y1) {
    x0 = max(0, __double2int_rd(__dsub_rn(__dsub_rn(cx, kp.reach), 1.0)));
    x1 = min(__double2int_ru(__dadd_rn(__dadd_rn(cx, kp.reach), 1.0)), kp.mod_w - 1);
    y0 = max(0, __double2int_rd(__dsub_rn(__dsub_rn(cy, kp.reach), 1.0)));
    y1 = min(__double2int_ru(__dadd_rn(__dadd_rn(cy, kp.reach), 1.0)), kp.mod_h - 1);
}

// ------------------------------------------------------------------------------------------------------------
// FP64 evaluation of one canvas pixel over the cells [ix0,ix1] x [jy0,jy1] (also the precision fallback of the
// FP32 kernel).
// ------------------------------------------------------------------------------------------------------------
template <typename TI, int NC>
__device__ __forceinline__ void pixel_f64(const AaiKernelParams &kp, double cx, double cy, int ix0, int ix1, int jy0,
                                          int jy1, double &sumA, double (&acc)[NC]) {
    sumA = 0.0;
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) acc[ch] = 0.0;
    if (ix0 > ix1 || jy0 > jy1) return;
    const AaiShape &g = kp.shape;
    // chord of the footprint on the horizontal grid line through the top of row jy0 (relative to C)
    double xlT, xrT;
    aai_chord_h(g, ((double)jy0 - 0.5) - cy, xlT, xrT);
    for (int j = jy0; j <= jy1; ++j) {
        const double ry = (double)j - cy;
        double xlB, xrB;
        aai_chord_h(g, ry + 0.5, xlB, xrB);
        // chord on the vertical grid line through the left of column ix0
        double yt, yb;
        aai_chord_v(g, ((double)ix0 - 0.5) - cx, yt, yb);
        double lenL = aai_overlap1(yt, yb, ry);
        for (int i = ix0; i <= ix1; ++i) {
            const double rx = (double)i - cx;
            aai_chord_v(g, rx + 0.5, yt, yb);
            const double lenR = aai_overlap1(yt, yb, ry);
            const double lenT = aai_overlap1(xlT, xrT, rx);
            const double lenB = aai_overlap1(xlB, xrB, rx);
            const double area = aai_cell_area(g, rx, ry, lenT, lenB, lenL, lenR);
            lenL = lenR;
            if (area != 0.0) {
                int sx, sy;
                mod_to_src(kp, i, j, sx, sy);
                const char *row = (const char *)kp.src + (int64_t)(sy - kp.src_y0) * kp.src_pitch;
                sumA += area;
#pragma unroll
                for (int ch = 0; ch < NC; ++ch) acc[ch] += SrcLoad<TI>::get(row, sx * NC + ch) * area;
            }
        }
        xlT = xlB;
        xrT = xrB;
    }
}

// cells that can have non-zero overlap: |i - cx| < hb + 1/2, intersected with the reference's clamped window
__device__ __forceinline__ void cell_range(const AaiKernelParams &kp, double cx, double cy, int &ix0, int &ix1,
                                           int &jy0, int &jy1) {
    int wx0, wx1, wy0, wy1;
    search_window(kp, cx, cy, wx0, wx1, wy0, wy1);
    const double ext = kp.hb + 0.5 + 1e-9;
    ix0 = max(wx0, __double2int_ru(cx - ext));
    ix1 = min(wx1, __double2int_rd(cx + ext));
    jy0 = max(wy0, __double2int_ru(cy - ext));
    jy1 = min(wy1, __double2int_rd(cy + ext));
}

// ------------------------------------------------------------------------------------------------------------
// Overlap kernel, FP64 arithmetic: one thread per canvas pixel, global (read-only path) source loads.
// ------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int NC>
__global__ void __launch_bounds__(TILE_W *TILE_H)
    overlap_kernel_f64(const __grid_constant__ AaiKernelParams kp) {
    const int x = blockIdx.x * TILE_W + threadIdx.x;
    const int y = kp.row0 + blockIdx.y * TILE_H + threadIdx.y;
    if (x >= kp.dst_w || y >= kp.row1) return;
    double cx, cy;
    pixel_centre(kp, x, y, cx, cy);
    int ix0, ix1, jy0, jy1;
    cell_range(kp, cx, cy, ix0, ix1, jy0, jy1);
    double sumA, acc[NC];
    pixel_f64<TI, NC>(kp, cx, cy, ix0, ix1, jy0, jy1, sumA, acc);
    char *drow = (char *)kp.dst + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
    const bool ok = DBL_EPSILON < fabs(sumA);  // Source.cpp:577
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, ok ? acc[ch] / sumA : 0.0);
}

// ------------------------------------------------------------------------------------------------------------
// Overlap kernel, FP32 arithmetic (the north star's "FP32 kernel"): same formulation on footprint-local FP32
// coordinates; at most MAXN x MAXN cells per pixel, column loop unrolled so that the MAXN+1 vertical-line chords
// live in registers; side lengths via FADD.SAT instead of min/max.  Pixels whose quirk decision falls inside the
// FP32 guard band, or whose total overlap is tiny (border slivers), are redone in FP64 by pixel_f64().
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct SrcLoadF;
template <>
struct SrcLoadF<double> {
    static __device__ __forceinline__ float get(const void *row, int idx) { return (float)__ldg((const double *)row + idx); }
};
template <>
struct SrcLoadF<float> {
    static __device__ __forceinline__ float get(const void *row, int idx) { return __ldg((const float *)row + idx); }
};
template <>
struct SrcLoadF<uint8_t> {
    static __device__ __forceinline__ float get(const void *row, int idx) {
        return (float)__ldg((const uint8_t *)row + idx);
    }
};

template <typename TI, typename TO, int NC, int MAXN>
__global__ void __launch_bounds__(TILE_W *TILE_H)
    overlap_kernel_f32(const __grid_constant__ AaiKernelParams kp) {
    const int x = blockIdx.x * TILE_W + threadIdx.x;
    const int y = kp.row0 + blockIdx.y * TILE_H + threadIdx.y;
    if (x >= kp.dst_w || y >= kp.row1) return;
    double cx, cy;
    pixel_centre(kp, x, y, cx, cy);
    int ix0, ix1, jy0, jy1;
    cell_range(kp, cx, cy, ix0, ix1, jy0, jy1);
    const int ncols = ix1 - ix0 + 1, nrows = jy1 - jy0 + 1;
    char *drow = (char *)kp.dst + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
    if (ncols <= 0 || nrows <= 0) {  // footprint bounding box misses the image: the reference writes 0 (577)
#pragma unroll
        for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, 0.0);
        return;
    }
    float sumA = 0.0f, acc[NC];
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) acc[ch] = 0.0f;
    bool redo = ncols > MAXN || nrows > MAXN;  // cannot happen for the MAXN the host picked; stay correct anyway
    if (!redo) {
        const AaiShapeF &g = kp.shapef;
        const double rcx = rint(cx), rcy = rint(cy);
        const float fx = (float)(cx - rcx), fy = (float)(cy - rcy);
        const int dj0 = jy0 - (int)rcy;
        const float rx0 = (float)(ix0 - (int)rcx) - fx;
        float yt[MAXN + 1], yb[MAXN + 1];
#pragma unroll
        for (int k = 0; k <= MAXN; ++k) aai_chord_v_f32(g, rx0 + ((float)k - 0.5f), yt[k], yb[k]);
        float xlT, xrT;
        aai_chord_h_f32(g, ((float)dj0 - fy) - 0.5f, xlT, xrT);
#pragma unroll 1
        for (int r = 0; r < nrows; ++r) {
            const float ry = (float)(dj0 + r) - fy;
            float xlB, xrB;
            aai_chord_h_f32(g, ry + 0.5f, xlB, xrB);
            const float ey = ry - 0.5f;
            float lenL = aai_overlap1_f32(yt[0], yb[0], ey);
            const float ur = -ry * g.sn, vr = ry * g.cs;
            const int j = jy0 + r;
#pragma unroll
            for (int k = 0; k < MAXN; ++k) {
                const float rx = rx0 + (float)k;
                const float ex = rx - 0.5f;
                const float lenR = aai_overlap1_f32(yt[k + 1], yb[k + 1], ey);
                const float lenT = aai_overlap1_f32(xlT, xrT, ex);
                const float lenB = aai_overlap1_f32(xlB, xrB, ex);
                const float u0 = fmaf(rx, g.cs, ur), v0 = fmaf(rx, g.sn, vr);
                float area = aai_cell_area_f32(g, u0, v0, lenT, lenB, lenL, lenR, redo);
                lenL = lenR;
                area = (k < ncols) ? area : 0.0f;
                int sx, sy;
                mod_to_src(kp, ix0 + min(k, ncols - 1), j, sx, sy);
                const char *row = (const char *)kp.src + (int64_t)(sy - kp.src_y0) * kp.src_pitch;
                sumA += area;
#pragma unroll
                for (int ch = 0; ch < NC; ++ch) acc[ch] = fmaf(SrcLoadF<TI>::get(row, sx * NC + ch), area, acc[ch]);
            }
            xlT = xlB;
            xrT = xrB;
        }
        redo = redo || sumA < 0.05f;  // border slivers (and FP32-invisible overlaps): keep relative accuracy
    }
    if (redo) {
        double s64, a64[NC];
        pixel_f64<TI, NC>(kp, cx, cy, ix0, ix1, jy0, jy1, s64, a64);
        const bool ok = DBL_EPSILON < fabs(s64);
#pragma unroll
        for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, ok ? a64[ch] / s64 : 0.0);
    } else {
        const float inv = 1.0f / sumA;
#pragma unroll
        for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, (double)(acc[ch] * inv));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Separable kernel v1 (reduced angle exactly 0): footprint = axis-aligned square, overlap = wx(i) * wy(j).
// ------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int NC>
__global__ void __launch_bounds__(TILE_W *TILE_H)
    separable_kernel_f64(const __grid_constant__ AaiKernelParams kp) {
    const int x = blockIdx.x * TILE_W + threadIdx.x;
    const int y = kp.row0 + blockIdx.y * TILE_H + threadIdx.y;
    if (x >= kp.dst_w || y >= kp.row1) return;
    double cx, cy;
    pixel_centre(kp, x, y, cx, cy);
    int wx0, wx1, wy0, wy1;
    search_window(kp, cx, cy, wx0, wx1, wy0, wy1);
    const double ext = kp.shape.half + 0.5 + 1e-9;
    const int ix0 = max(wx0, __double2int_ru(cx - ext)), ix1 = min(wx1, __double2int_rd(cx + ext));
    const int jy0 = max(wy0, __double2int_ru(cy - ext)), jy1 = min(wy1, __double2int_rd(cy + ext));
    const double xl = cx - kp.shape.half, xr = cx + kp.shape.half, yt = cy - kp.shape.half, yb = cy + kp.shape.half;
    double sumA = 0.0;
    double acc[NC];
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) acc[ch] = 0.0;
    for (int j = jy0; j <= jy1; ++j) {
        const double wy = fmax(fmin(yb, (double)j + 0.5) - fmax(yt, (double)j - 0.5), 0.0);
        for (int i = ix0; i <= ix1; ++i) {
            const double wx = fmax(fmin(xr, (double)i + 0.5) - fmax(xl, (double)i - 0.5), 0.0);
            const double area = wx * wy;
            if (area != 0.0) {
                int sx, sy;
                mod_to_src(kp, i, j, sx, sy);
                const char *row = (const char *)kp.src + (int64_t)(sy - kp.src_y0) * kp.src_pitch;
                sumA += area;
#pragma unroll
                for (int ch = 0; ch < NC; ++ch) acc[ch] += SrcLoad<TI>::get(row, sx * NC + ch) * area;
            }
        }
    }
    char *drow = (char *)kp.dst + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
    const bool ok = DBL_EPSILON < fabs(sumA);
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, ok ? acc[ch] / sumA : 0.0);
}

// ------------------------------------------------------------------------------------------------------------
// Fast mode (Source.cpp:866-907): unweighted mean of the expanded pixels whose CENTRE lies in the footprint.
// ------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int NC>
__global__ void __launch_bounds__(TILE_W *TILE_H) fast_kernel(const __grid_constant__ AaiKernelParams kp) {
    const int x = blockIdx.x * TILE_W + threadIdx.x;
    const int y = kp.row0 + blockIdx.y * TILE_H + threadIdx.y;
    if (x >= kp.dst_w || y >= kp.row1) return;
    double cx, cy;
    pixel_centre(kp, x, y, cx, cy);
    int wx0, wx1, wy0, wy1;
    search_window(kp, cx, cy, wx0, wx1, wy0, wy1);
    const double ext = kp.hb + 1e-6;
    const int ix0 = max(wx0, __double2int_ru(cx - ext)), ix1 = min(wx1, __double2int_rd(cx + ext));
    const int jy0 = max(wy0, __double2int_ru(cy - ext)), jy1 = min(wy1, __double2int_rd(cy + ext));
    int count = 0;
    double acc[NC];
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) acc[ch] = 0.0;
    for (int j = jy0; j <= jy1; ++j) {
        const double ry = (double)j - cy;
        for (int i = ix0; i <= ix1; ++i) {
            const double rx = (double)i - cx;
            const double u0 = rx * kp.shape.cs - ry * kp.shape.sn;
            const double v0 = rx * kp.shape.sn + ry * kp.shape.cs;
            if (fabs(u0) <= kp.shape.half && fabs(v0) <= kp.shape.half) {  // closed point-in-square (837-864)
                int sx, sy;
                mod_to_src(kp, i, j, sx, sy);
                const char *row = (const char *)kp.src + (int64_t)(sy - kp.src_y0) * kp.src_pitch;
                count += 1;
#pragma unroll
                for (int ch = 0; ch < NC; ++ch) acc[ch] += SrcLoad<TI>::get(row, sx * NC + ch);
            }
        }
    }
    char *drow = (char *)kp.dst + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, count > 0 ? acc[ch] / (double)count : 0.0);
}

enum KernelKind { K_OVERLAP, K_OVERLAP_F32_4, K_OVERLAP_F32_5, K_OVERLAP_F32_6, K_OVERLAP_F32_8, K_SEPARABLE, K_FAST };

template <typename TI, typename TO, int NC>
cudaError_t launch_typed(KernelKind kind, const AaiKernelParams &kp, cudaStream_t stream) {
    const int rows = kp.row1 - kp.row0;
    if (rows <= 0 || kp.dst_w <= 0) return cudaSuccess;
    dim3 block(TILE_W, TILE_H);
    dim3 grid((kp.dst_w + TILE_W - 1) / TILE_W, (rows + TILE_H - 1) / TILE_H);
    switch (kind) {
        case K_OVERLAP: overlap_kernel_f64<TI, TO, NC><<<grid, block, 0, stream>>>(kp); break;
        case K_OVERLAP_F32_4: overlap_kernel_f32<TI, TO, NC, 4><<<grid, block, 0, stream>>>(kp); break;
        case K_OVERLAP_F32_5: overlap_kernel_f32<TI, TO, NC, 5><<<grid, block, 0, stream>>>(kp); break;
        case K_OVERLAP_F32_6: overlap_kernel_f32<TI, TO, NC, 6><<<grid, block, 0, stream>>>(kp); break;
        case K_OVERLAP_F32_8: overlap_kernel_f32<TI, TO, NC, 8><<<grid, block, 0, stream>>>(kp); break;
        case K_SEPARABLE: separable_kernel_f64<TI, TO, NC><<<grid, block, 0, stream>>>(kp); break;
        case K_FAST: fast_kernel<TI, TO, NC><<<grid, block, 0, stream>>>(kp); break;
    }
    return cudaGetLastError();
}

template <typename TI, typename TO>
cudaError_t launch_channels(KernelKind kind, const AaiKernelParams &kp, cudaStream_t stream) {
    switch (kp.channels) {
        case 1: return launch_typed<TI, TO, 1>(kind, kp, stream);
        case 2: return launch_typed<TI, TO, 2>(kind, kp, stream);
        case 3: return launch_typed<TI, TO, 3>(kind, kp, stream);
        case 4: return launch_typed<TI, TO, 4>(kind, kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <typename TI>
cudaError_t launch_dst(KernelKind kind, const AaiKernelParams &kp, int dst_dtype, cudaStream_t stream) {
    switch (dst_dtype) {
        case AAI_F64: return launch_channels<TI, double>(kind, kp, stream);
        case AAI_F32: return launch_channels<TI, float>(kind, kp, stream);
        case AAI_U8: return launch_channels<TI, uint8_t>(kind, kp, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_any(KernelKind kind, const AaiKernelParams &kp, int src_dtype, int dst_dtype, cudaStream_t stream) {
    switch (src_dtype) {
        case AAI_F64: return launch_dst<double>(kind, kp, dst_dtype, stream);
        case AAI_F32: return launch_dst<float>(kind, kp, dst_dtype, stream);
        case AAI_U8: return launch_dst<uint8_t>(kind, kp, dst_dtype, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

int aai_launch_overlap(const AaiKernelParams &kp, int arith, int src_dtype, int dst_dtype, void *stream) {
    KernelKind kind = K_OVERLAP;
    if (arith == AAI_ARITH_F32 && kp.f32_ok) {
        // at most floor(2*hb + 1) + 1 cells per axis have non-zero overlap
        const int n = (int)floor(2.0 * kp.hb + 1.0 + 2e-9) + 1;
        if (n <= 4) kind = K_OVERLAP_F32_4;
        else if (n <= 5) kind = K_OVERLAP_F32_5;
        else if (n <= 6) kind = K_OVERLAP_F32_6;
        else if (n <= 8) kind = K_OVERLAP_F32_8;
    }
    return (int)launch_any(kind, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
int aai_launch_separable(const AaiKernelParams &kp, int arith, int src_dtype, int dst_dtype, void *stream) {
    (void)arith;
    return (int)launch_any(K_SEPARABLE, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
int aai_launch_fast(const AaiKernelParams &kp, int src_dtype, int dst_dtype, void *stream) {
    return (int)launch_any(K_FAST, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
