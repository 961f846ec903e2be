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
// Overlap kernel v1: one thread per canvas pixel, FP64, global (read-only path) source loads.
// ------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int NC>
__global__ void __launch_bounds__(TILE_W *TILE_H)
    overlap_kernel_f64(const __grid_constant__ AaiKernelParams kp) {
    const int x = blockIdx.x * TILE_W + threadIdx.x;
    const int y = kp.row0 + blockIdx.y * TILE_H + threadIdx.y;
    if (x >= kp.dst_w || y >= kp.row1) return;

    double cx, cy;
    pixel_centre(kp, x, y, cx, cy);
    int wx0, wx1, wy0, wy1;
    search_window(kp, cx, cy, wx0, wx1, wy0, wy1);
    // cells that can have non-zero overlap: |i - cx| < hb + 1/2 (a subset of the reference's window)
    const double ext = kp.hb + 0.5 + 1e-9;
    const int ix0 = max(wx0, __double2int_ru(cx - ext)), ix1 = min(wx1, __double2int_rd(cx + ext));
    const int jy0 = max(wy0, __double2int_ru(cy - ext)), jy1 = min(wy1, __double2int_rd(cy + ext));

    double sumA = 0.0;
    double acc[NC];
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) acc[ch] = 0.0;

    if (ix0 <= ix1 && jy0 <= jy1) {
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
    char *drow = (char *)kp.dst + (int64_t)(y - kp.dst_y0) * kp.dst_pitch;
    const bool ok = DBL_EPSILON < fabs(sumA);  // Source.cpp:577
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) store_dst<TO>(drow, x * NC + ch, ok ? acc[ch] / sumA : 0.0);
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

enum KernelKind { K_OVERLAP, K_SEPARABLE, K_FAST };

template <typename TI, typename TO, int NC>
cudaError_t launch_typed(KernelKind kind, const AaiKernelParams &kp, cudaStream_t stream) {
    const int rows = kp.row1 - kp.row0;
    if (rows <= 0 || kp.dst_w <= 0) return cudaSuccess;
    dim3 block(TILE_W, TILE_H);
    dim3 grid((kp.dst_w + TILE_W - 1) / TILE_W, (rows + TILE_H - 1) / TILE_H);
    switch (kind) {
        case K_OVERLAP: overlap_kernel_f64<TI, TO, NC><<<grid, block, 0, stream>>>(kp); break;
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
    (void)arith;
    return (int)launch_any(K_OVERLAP, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
int aai_launch_separable(const AaiKernelParams &kp, int arith, int src_dtype, int dst_dtype, void *stream) {
    (void)arith;
    return (int)launch_any(K_SEPARABLE, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
int aai_launch_fast(const AaiKernelParams &kp, int src_dtype, int dst_dtype, void *stream) {
    return (int)launch_any(K_FAST, kp, src_dtype, dst_dtype, (cudaStream_t)stream);
}
