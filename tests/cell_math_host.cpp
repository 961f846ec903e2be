// TEST-ONLY host build of the device cell arithmetic (csrc/aai_cell.cuh), so that the CPU suite can compare
// the kernels' closed form with the oracle pair by pair without a GPU.  Never linked into the product.
#include <cmath>
#include "cell_legacy_forms.h"

static AaiShape make_shape(double c, double s, double L) { return aai_make_shape(c, s, L); }

extern "C" void aai_test_pair_areas(double c, double s, double L, const double *cx, const double *cy, const int *i,
                                    const int *j, double *out, long long n) {
    const AaiShape g = make_shape(c, s, L);
    for (long long k = 0; k < n; ++k) out[k] = aai_pair_area(g, cx[k], cy[k], i[k], j[k]);
}

// Whole-image evaluation with the device arithmetic (FP64 kernel logic restated for the host), TEST ONLY.
extern "C" void aai_test_image_f64(double c, double s, double L, double offIx, double offIy, double isoX, double isoY,
                                   double offX, double offY, int modW, int modH, int dstW, int dstH,
                                   const double *mod /* modH x modW */, double *out) {
    const AaiShape g = make_shape(c, s, L);
    const double reach = L * std::sqrt(2) / 2, hb = g.half * (c + s), ext = hb + 0.5 + 1e-9;
    for (int y = 0; y < dstH; ++y)
        for (int x = 0; x < dstW; ++x) {
            const double u = ((x + offIx) * L - isoX) + offX, v = ((y + offIy) * L - isoY) + offY;
            const double cx = (u * c + v * s) + isoX, cy = (-u * s + v * c) + isoY;
            int wx0 = std::max(0, (int)std::floor(cx - reach - 1)), wx1 = std::min((int)std::ceil(cx + reach + 1), modW - 1);
            int wy0 = std::max(0, (int)std::floor(cy - reach - 1)), wy1 = std::min((int)std::ceil(cy + reach + 1), modH - 1);
            int ix0 = std::max(wx0, (int)std::ceil(cx - ext)), ix1 = std::min(wx1, (int)std::floor(cx + ext));
            int jy0 = std::max(wy0, (int)std::ceil(cy - ext)), jy1 = std::min(wy1, (int)std::floor(cy + ext));
            double sumA = 0, acc = 0;
            for (int j = jy0; j <= jy1; ++j)
                for (int i = ix0; i <= ix1; ++i) {
                    const double a = aai_pair_area(g, cx, cy, i, j);
                    sumA += a;
                    acc += mod[(size_t)j * modW + i] * a;
                }
            out[(size_t)y * dstW + x] = 2.220446049250313e-16 < std::fabs(sumA) ? acc / sumA : 0.0;
        }
}

static AaiShapeF make_shape_f(double c, double s, double L) { return aai_make_shape_f(c, s, L); }

// FP32 pair areas + the "uncertain" flag (1 = the kernel would redo this pixel in FP64)
extern "C" void aai_test_pair_areas_f32(double c, double s, double L, const double *cx, const double *cy, const int *i,
                                        const int *j, float *out, unsigned char *flag, long long n) {
    const AaiShapeF g = make_shape_f(c, s, L);
    for (long long k = 0; k < n; ++k) {
        const double ix = std::nearbyint(cx[k]), iy = std::nearbyint(cy[k]);
        float worst = 1.0f;
        out[k] = aai_pair_area_f32(g, (float)(cx[k] - ix), (float)(cy[k] - iy), i[k] - (int)ix, j[k] - (int)iy, worst);
        flag[k] = worst < g.tau ? 1 : 0;
    }
}

// Whole-image evaluation with the FP32 kernel's arithmetic (no FP64 fallback: flagged pixels are reported), TEST ONLY.
extern "C" void aai_test_image_f32(double c, double s, double L, double offIx, double offIy, double isoX, double isoY,
                                   double offX, double offY, int modW, int modH, int dstW, int dstH,
                                   const double *mod /* modH x modW */, double *out, unsigned char *flag) {
    const AaiShapeF g = make_shape_f(c, s, L);
    const double reach = L * std::sqrt(2) / 2, hb = (L / 2) * (c + s), ext = hb + 0.5 + 1e-9;
    for (int y = 0; y < dstH; ++y)
        for (int x = 0; x < dstW; ++x) {
            const double u = ((x + offIx) * L - isoX) + offX, v = ((y + offIy) * L - isoY) + offY;
            const double cx = (u * c + v * s) + isoX, cy = (-u * s + v * c) + isoY;
            int wx0 = std::max(0, (int)std::floor(cx - reach - 1)), wx1 = std::min((int)std::ceil(cx + reach + 1), modW - 1);
            int wy0 = std::max(0, (int)std::floor(cy - reach - 1)), wy1 = std::min((int)std::ceil(cy + reach + 1), modH - 1);
            int ix0 = std::max(wx0, (int)std::ceil(cx - ext)), ix1 = std::min(wx1, (int)std::floor(cx + ext));
            int jy0 = std::max(wy0, (int)std::ceil(cy - ext)), jy1 = std::min(wy1, (int)std::floor(cy + ext));
            const double rcx = std::nearbyint(cx), rcy = std::nearbyint(cy);
            const float fx = (float)(cx - rcx), fy = (float)(cy - rcy);
            float sumA = 0, acc = 0, worst = 1.0f;
            for (int j = jy0; j <= jy1; ++j)
                for (int i = ix0; i <= ix1; ++i) {
                    const float a = aai_pair_area_f32(g, fx, fy, i - (int)rcx, j - (int)rcy, worst);
                    sumA += a;
                    acc = fmaf((float)mod[(size_t)j * modW + i], a, acc);
                }
            const size_t k = (size_t)y * dstW + x;
            flag[k] = (worst < g.tau || sumA < 0.05f) ? 1 : 0;
            out[k] = sumA > 0 ? (double)(acc * (1.0f / sumA)) : 0.0;
        }
}


// Packed (two cells per call) variant: evaluates cells (i, j) and (i+1, j); returns lane .x in out0, lane .y in out1.
extern "C" void aai_test_pair_areas_f32x2(double c, double s, double L, const double *cx, const double *cy, const int *i,
                                          const int *j, float *out0, float *out1, unsigned char *flag, long long n) {
    const AaiShapeF g = make_shape_f(c, s, L);
    for (long long k = 0; k < n; ++k) {
        const double ix = std::nearbyint(cx[k]), iy = std::nearbyint(cy[k]);
        const float fx = (float)(cx[k] - ix), fy = (float)(cy[k] - iy);
        const int di = i[k] - (int)ix, dj = j[k] - (int)iy;
        const float rx = (float)di - fx, ry = (float)dj - fy;
        float xlT, xrT, xlB, xrB, y0t, y0b, y1t, y1b, y2t, y2b;
        aai_chord_h_f32(g, ry - 0.5f, xlT, xrT);
        aai_chord_h_f32(g, ry + 0.5f, xlB, xrB);
        aai_chord_v_f32(g, rx - 0.5f, y0t, y0b);
        aai_chord_v_f32(g, rx + 0.5f, y1t, y1b);
        aai_chord_v_f32(g, rx + 1.5f, y2t, y2b);
        const float ey = ry - 0.5f;
        const float l0 = aai_overlap1_f32(y0t, y0b, ey), l1 = aai_overlap1_f32(y1t, y1b, ey), l2 = aai_overlap1_f32(y2t, y2b, ey);
        const AaiF2 lenT = aai_f2(aai_overlap1_f32(xlT, xrT, rx - 0.5f), aai_overlap1_f32(xlT, xrT, rx + 0.5f));
        const AaiF2 lenB = aai_f2(aai_overlap1_f32(xlB, xrB, rx - 0.5f), aai_overlap1_f32(xlB, xrB, rx + 0.5f));
        const AaiF2 u0 = aai_f2(fmaf(rx, g.cs, -ry * g.sn), fmaf(rx + 1.0f, g.cs, -ry * g.sn));
        const AaiF2 v0 = aai_f2(fmaf(rx, g.sn, ry * g.cs), fmaf(rx + 1.0f, g.sn, ry * g.cs));
        float worst = 1.0f;
        const AaiF2 a = aai_cell_area_f32x2(g, u0, v0, lenT, lenB, aai_f2(l0, l1), aai_f2(l1, l2), worst);
        out0[k] = a.x;
        out1[k] = a.y;
        flag[k] = worst < g.tau ? 1 : 0;
    }
}


// Per-cell areas of one footprint over an n x n block of cells starting at (i0, j0), evaluated the way the FP32 kernel
// does it (exact Green areas + per-row quirk events).  out: n*n floats (row-major), returns the decision margin.
extern "C" float aai_test_footprint_rows_f32(double c, double s, double L, double cx, double cy, int i0, int j0, int n,
                                             float *out) {
    const AaiShapeF g = make_shape_f(c, s, L);
    const double rcx = std::nearbyint(cx), rcy = std::nearbyint(cy);
    const float fx = (float)(cx - rcx), fy = (float)(cy - rcy);
    const float rx0 = (float)(i0 - (int)rcx) - fx;
    const int dj0 = j0 - (int)rcy;
    float worst = 1.0f;
    for (int r = 0; r < n; ++r) {
        const float ry = (float)(dj0 + r) - fy;
        float xlT, xrT, lineL, lineR, xlB, xrB;
        aai_chord_h_f32(g, ry - 0.5f, xlT, xrT, lineL, lineR);
        aai_chord_h_f32(g, ry + 0.5f, xlB, xrB);
        const float ey = ry - 0.5f, ur = -ry * g.sn, vr = ry * g.cs;
        for (int k = 0; k < n; ++k) {
            const float rx = rx0 + (float)k, ex = rx - 0.5f;
            float ytL, ybL, ytR, ybR;
            aai_chord_v_f32(g, rx - 0.5f, ytL, ybL);
            aai_chord_v_f32(g, rx + 0.5f, ytR, ybR);
            out[r * n + k] = aai_cell_exact_f32(g, fmaf(rx, g.cs, ur), fmaf(rx, g.sn, vr), aai_overlap1_f32(xlT, xrT, ex),
                                                aai_overlap1_f32(xlB, xrB, ex), aai_overlap1_f32(ytL, ybL, ey),
                                                aai_overlap1_f32(ytR, ybR, ey));
        }
        const float e0 = rx0 - 0.5f;
        int ka, kb;
        float da, db;
        aai_row_quirk_f32<true>(g, lineL - e0, ey, rx0, vr, ka, da, kb, db, worst);
        if (ka >= 0 && ka < n) out[r * n + ka] += da;
        if (kb >= 0 && kb < n) out[r * n + kb] += db;
        aai_row_quirk_f32<false>(g, lineR - e0, ey, rx0, vr, ka, da, kb, db, worst);
        if (ka >= 0 && ka < n) out[r * n + ka] += da;
        if (kb >= 0 && kb < n) out[r * n + kb] += db;
    }
    return worst;
}


// Same block of cells evaluated the way the FP32 kernel does it NOW: exact Green areas + per-edge quirk events
// (aai_edge_quirk_f32).  out: n*n floats (row-major); *total = L^2 + sum of the corrections (the kernel's sumA);
// returns the decision margin.
extern "C" float aai_test_footprint_edges_f32(double c, double s, double L, double cx, double cy, int i0, int j0, int n,
                                              float *out, float *total) {
    const AaiShapeF g = make_shape_f(c, s, L);
    const double rcx = std::nearbyint(cx), rcy = std::nearbyint(cy);
    const float fx = (float)(cx - rcx), fy = (float)(cy - rcy);
    const float rx0 = (float)(i0 - (int)rcx) - fx;
    const int dj0 = j0 - (int)rcy;
    for (int r = 0; r < n; ++r) {
        const float ry = (float)(dj0 + r) - fy;
        float xlT, xrT, xlB, xrB;
        aai_chord_h_f32(g, ry - 0.5f, xlT, xrT);
        aai_chord_h_f32(g, ry + 0.5f, xlB, xrB);
        const float ey = ry - 0.5f, ur = -ry * g.sn, vr = ry * g.cs;
        for (int k = 0; k < n; ++k) {
            const float rx = rx0 + (float)k, ex = rx - 0.5f;
            float ytL, ybL, ytR, ybR;
            aai_chord_v_f32(g, rx - 0.5f, ytL, ybL);
            aai_chord_v_f32(g, rx + 0.5f, ytR, ybR);
            out[r * n + k] = aai_cell_exact_f32(g, fmaf(rx, g.cs, ur), fmaf(rx, g.sn, vr), aai_overlap1_f32(xlT, xrT, ex),
                                                aai_overlap1_f32(xlB, xrB, ex), aai_overlap1_f32(ytL, ybL, ey),
                                                aai_overlap1_f32(ytR, ybR, ey));
        }
    }
    const float e0 = rx0 - 0.5f, t0 = ((float)dj0 - fy) - 0.5f;
    const float g0m = g.steep ? e0 : t0, g0M = g.steep ? t0 : e0;
    float worst = 1.0f, sum = g.area_total;
    auto apply = [&](int mi, int Mi, float d) {
        const int k = g.steep ? mi : Mi, r = g.steep ? Mi : mi;
        if (d != 0.0f && k >= 0 && k < n && r >= 0 && r < n) {
            out[r * n + k] += d;
            sum += d;
        }
    };
    for (int q = 0; q < g.ncross; ++q) {
        int mi, Mi;
        float db, da;
        aai_edge_quirk_f32<true>(g, g0m, g0M, q, mi, Mi, db, da, worst);
        apply(mi, Mi, db);
        apply(mi + 1, Mi, da);
        aai_edge_quirk_f32<false>(g, g0m, g0M, q, mi, Mi, db, da, worst);
        apply(mi, Mi, db);
        apply(mi + 1, Mi, da);
    }
    *total = sum;
    return worst;
}


// The packed two-edge routine against the scalar one: returns the number of mismatching outputs over n random grids.
extern "C" long long aai_test_edge_pair_vs_scalar(double c, double s, double L, const double *g0m, const double *g0M,
                                                  long long n) {
    const AaiShapeF g = make_shape_f(c, s, L);
    long long bad = 0;
    for (long long k = 0; k < n; ++k)
        for (int q = 0; q < g.ncross; ++q) {
            int mi[2], Mi[2], m0, M0, m1, M1;
            float db[2], da[2], b0, a0, b1, a1, w = 1.0f, w0 = 1.0f;
            aai_edge_quirk_pair_f32(g, (float)g0m[k], (float)g0M[k], q, mi, Mi, db, da, w);
            aai_edge_quirk_f32<true>(g, (float)g0m[k], (float)g0M[k], q, m0, M0, b0, a0, w0);
            aai_edge_quirk_f32<false>(g, (float)g0m[k], (float)g0M[k], q, m1, M1, b1, a1, w0);
            if (mi[0] != m0 || Mi[0] != M0 || mi[1] != m1 || Mi[1] != M1 || db[0] != b0 || da[0] != a0 || db[1] != b1 ||
                da[1] != a1 || w != w0)
                ++bad;
        }
    return bad;
}


// FP64 edge formulation (what the unrolled FP64 kernel runs): exact Green areas + per-edge events, n x n block.
extern "C" void aai_test_footprint_edges_f64(double c, double s, double L, double cx, double cy, int i0, int j0, int n,
                                             double *out, double *total) {
    const AaiShape g = make_shape(c, s, L);
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < n; ++k) {
            const double rx = (double)(i0 + k) - cx, ry = (double)(j0 + r) - cy;
            double xlT, xrT, xlB, xrB, ytL, ybL, ytR, ybR;
            aai_chord_h(g, ry - 0.5, xlT, xrT);
            aai_chord_h(g, ry + 0.5, xlB, xrB);
            aai_chord_v(g, rx - 0.5, ytL, ybL);
            aai_chord_v(g, rx + 0.5, ytR, ybR);
            // the kernel's arithmetic: side lengths as differences of clamped boundaries, folded Green form
            const double a = rx - 0.5, b = rx + 0.5, t = ry - 0.5, u = ry + 0.5;
            out[r * n + k] = aai_cell_exact_f64(g, rx, ry, aai_clamp_chord(b, xlT, xrT) - aai_clamp_chord(a, xlT, xrT),
                                                aai_clamp_chord(b, xlB, xrB) - aai_clamp_chord(a, xlB, xrB),
                                                aai_clamp_chord(u, ytL, ybL) - aai_clamp_chord(t, ytL, ybL),
                                                aai_clamp_chord(u, ytR, ybR) - aai_clamp_chord(t, ytR, ybR));
        }
    const double e0 = ((double)i0 - cx) - 0.5, t0 = ((double)j0 - cy) - 0.5;
    const double g0m = g.steep ? e0 : t0, g0M = g.steep ? t0 : e0;
    double sum = g.area_total;
    auto apply = [&](int mi, int Mi, double d) {
        const int k = g.steep ? mi : Mi, r = g.steep ? Mi : mi;
        if (d != 0.0 && k >= 0 && k < n && r >= 0 && r < n) {
            out[r * n + k] += d;
            sum += d;
        }
    };
    for (int q = 0; q < g.ncross; ++q) {
        int mi, Mi;
        double db, da;
        aai_edge_quirk_f64<true>(g, g0m, g0M, q, mi, Mi, db, da);
        apply(mi, Mi, db);
        apply(mi + 1, Mi, da);
        aai_edge_quirk_f64<false>(g, g0m, g0M, q, mi, Mi, db, da);
        apply(mi, Mi, db);
        apply(mi + 1, Mi, da);
    }
    *total = sum;
}

// Quadrant weights of the upscaling path (aai_quadrant_areas_f32) against the sum of the exact per-cell areas (FP64
// Green form, quirk off) of the cells on each side of the source-pixel boundaries.  (cx, cy): footprint centre in the
// expanded frame; boundaries at multiples of S minus 1/2.  out[k] = {W00, W01, W10, W11, R00, R01, R10, R11}.
extern "C" void aai_test_quadrant_areas(double c, double s, double L, int S, const double *cx, const double *cy,
                                        double *out, long long n) {
    const AaiShape g = make_shape(c, s, L);
    const AaiShapeF gf = make_shape_f(c, s, L);
    const double ext = g.half * (c + s) + 0.5 + 1e-9;
    for (long long k = 0; k < n; ++k) {
        const int ix0 = (int)std::ceil(cx[k] - ext), ix1 = (int)std::floor(cx[k] + ext);
        const int jy0 = (int)std::ceil(cy[k] - ext), jy1 = (int)std::floor(cy[k] + ext);
        // first boundary to the right of cell ix0: after the cell i with (i + 1) % S == 0
        auto boundary = [&](int i0, int i1) -> double {  // coordinate of the boundary inside [i0, i1], or +1e6
            for (int i = i0; i < i1; ++i)
                if (((i + 1) % S + S) % S == 0) return (double)i + 0.5;
            return 1e6;
        };
        const double X = boundary(ix0, ix1), Y = boundary(jy0, jy1);
        const double irx = std::nearbyint(cx[k]), iry = std::nearbyint(cy[k]);
        const float fx = (float)(cx[k] - irx), fy = (float)(cy[k] - iry);
        const bool hasX = X < 1e5, hasY = Y < 1e5;
        const float tX = hasX ? (float)(X - irx) - fx : 0.0f, tY = hasY ? (float)(Y - iry) - fy : 0.0f;
        float w[4];
        aai_quadrant_areas_f32(gf, tX, tY, hasX, hasY, w[0], w[1], w[2], w[3]);
        double r[4] = {0, 0, 0, 0};
        for (int j = jy0; j <= jy1; ++j)
            for (int i = ix0; i <= ix1; ++i) {
                const double rx = i - cx[k], ry = j - cy[k];
                double xlT, xrT, xlB, xrB, ytL, ybL, ytR, ybR;
                aai_chord_h(g, ry - 0.5, xlT, xrT);
                aai_chord_h(g, ry + 0.5, xlB, xrB);
                aai_chord_v(g, rx - 0.5, ytL, ybL);
                aai_chord_v(g, rx + 0.5, ytR, ybR);
                const double a = aai_cell_area(g, rx, ry, aai_overlap1(xlT, xrT, rx), aai_overlap1(xlB, xrB, rx),
                                               aai_overlap1(ytL, ybL, ry), aai_overlap1(ytR, ybR, ry), false);
                r[(j > Y ? 2 : 0) + (i > X ? 1 : 0)] += a;
            }
        for (int q = 0; q < 4; ++q) {
            out[8 * k + q] = w[q];
            out[8 * k + 4 + q] = r[q];
        }
    }
}

// Whole-image evaluation with the arithmetic of the FP32 kernel's UPSCALING path (ADDR_GROUPED: quadrant weights +
// per-edge quirk events booked per source pixel), TEST ONLY.  `mod` is the expanded + quadrant-rotated source
// (modH x modW); the source-pixel groups along an axis are given by the expanded coordinate e = a * i + e0 (a = +-1)
// and the scale S.  flag: 1 = the kernel would redo the pixel in FP64 (guard band / border), out is then not filled.
extern "C" void aai_test_image_grouped_f32(double c, double s, double L, double offIx, double offIy, double isoX,
                                           double isoY, double offX, double offY, int modW, int modH, int dstW, int dstH,
                                           int S, int ac, int ec0, int ar, int er0, const double *mod, double *out,
                                           unsigned char *flag) {
    const AaiShapeF g = make_shape_f(c, s, L);
    const double h = L / 2;
    const float ext32 = (float)(h * (c + s) + 0.5 + 2e-6);
    auto first_group = [&](int e, int step) { const int rem = ((e % S) + S) % S; return step > 0 ? S - rem : rem + 1; };
    for (int y = 0; y < dstH; ++y)
        for (int x = 0; x < dstW; ++x) {
            const size_t idx = (size_t)y * dstW + x;
            const double u = ((x + offIx) * L - isoX) + offX, v = ((y + offIy) * L - isoY) + offY;
            const double cx = (u * c + v * s) + isoX, cy = (-u * s + v * c) + isoY;
            const int irx = (int)std::nearbyint(cx), iry = (int)std::nearbyint(cy);
            const float fx = (float)(cx - irx), fy = (float)(cy - iry);
            const int bx0 = irx + (int)std::ceil(fx - ext32), bx1 = irx + (int)std::floor(fx + ext32);
            const int by0 = iry + (int)std::ceil(fy - ext32), by1 = iry + (int)std::floor(fy + ext32);
            const int ix0 = std::max(0, bx0), ix1 = std::min(modW - 1, bx1), jy0 = std::max(0, by0), jy1 = std::min(modH - 1, by1);
            const int ncols = ix1 - ix0 + 1, nrows = jy1 - jy0 + 1;
            out[idx] = 0.0;
            flag[idx] = 0;
            if (ncols <= 0 || nrows <= 0) continue;
            if (bx0 < 0 || by0 < 0 || bx1 > modW - 1 || by1 > modH - 1 || ncols > 4 || nrows > 4) {
                flag[idx] = 1;
                continue;
            }
            const int dj0 = jy0 - iry;
            const float rx0 = (float)(ix0 - irx) - fx, t0 = ((float)dj0 - fy) - 0.5f, e0 = rx0 - 0.5f;
            const int nA = first_group(ac * ix0 + ec0, ac), nT = first_group(ar * jy0 + er0, ar);
            float W00, W01, W10, W11;
            aai_quadrant_areas_f32(g, e0 + (float)nA, t0 + (float)nT, nA < ncols, nT < nrows, W00, W01, W10, W11);
            float sumA = g.area_total, worst = 1.0f;
            const float g0m = g.steep ? e0 : t0, g0M = g.steep ? t0 : e0;
            for (int q = 0; q < g.ncross; ++q) {
                int mi[2], Mi[2];
                float db[2], da[2];
                aai_edge_quirk_pair_f32(g, g0m, g0M, q, mi, Mi, db, da, worst);
                for (int e = 0; e < 2; ++e) {
                    const int mlim = (g.steep ? ncols : nrows) - 2, Mlim = (g.steep ? nrows : ncols) - 1;
                    const int m = std::max(0, std::min(mi[e], mlim)), M = std::max(0, std::min(Mi[e], Mlim));
                    const int k = g.steep ? m : M, r = g.steep ? M : m;
                    const int k1 = k + (g.steep ? 1 : 0), r1 = r + (g.steep ? 0 : 1);
                    const float both = db[e] + da[e];
                    if (g.steep) {
                        const float dA = (k < nA ? db[e] : 0.0f) + (k1 < nA ? da[e] : 0.0f), dB = both - dA;
                        const bool top = r < nT;
                        W00 += top ? dA : 0.0f; W01 += top ? dB : 0.0f; W10 += top ? 0.0f : dA; W11 += top ? 0.0f : dB;
                    } else {
                        const float dT = (r < nT ? db[e] : 0.0f) + (r1 < nT ? da[e] : 0.0f), dBt = both - dT;
                        const bool left = k < nA;
                        W00 += left ? dT : 0.0f; W10 += left ? dBt : 0.0f; W01 += left ? 0.0f : dT; W11 += left ? 0.0f : dBt;
                    }
                    sumA += both;
                }
            }
            const float v00 = (float)mod[(size_t)jy0 * modW + ix0], v01 = (float)mod[(size_t)jy0 * modW + ix1];
            const float v10 = (float)mod[(size_t)jy1 * modW + ix0], v11 = (float)mod[(size_t)jy1 * modW + ix1];
            const float acc = fmaf(v00, W00, fmaf(v01, W01, fmaf(v10, W10, v11 * W11)));
            if (worst < g.tau || sumA < 0.25f) {
                flag[idx] = 1;
                continue;
            }
            out[idx] = (double)(acc * (1.0f / sumA));
        }
}
