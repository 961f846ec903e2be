// Host FP64 geometry plan + row-band partitioner of the area-average hot path.
//
// Replaces the scalar preamble of AreaAverageInterpolation::areaAverageInterpolation
// (Source.cpp:112-200): argument checks, integer expansion factor, quadrant reduction, canvas size,
// returned isocentre, canvas offset.  Integer sizes come out of round() / (int) truncation of FP64
// expressions, so every expression below keeps the reference's operand order; sin/cos are evaluated
// once here with the host libm, so that the device sees the constants the reference used.
//
// No CUDA in this file: it is exercised by the CPU test-suite.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "aai_internal.h"

namespace {

// canvas pixel centre in expanded-source coordinates; same expression as Source.cpp:212-219
inline void centre_of(const aai_plan &p, double x, double y, double &cx, double &cy) {
    const double u = (x + p.off_ix) * p.side - p.iso_x + p.off_x;
    const double v = (y + p.off_iy) * p.side - p.iso_y + p.off_y;
    cx = u * p.cos_t + v * p.sin_t + p.iso_x;
    cy = -u * p.sin_t + v * p.cos_t + p.iso_y;
}

// For canvas row y: the half-open x range whose search window (Source.cpp:426-429) can meet the source.
// The centre is affine in x, so each of the four window conditions is a half-line in x.
inline void covered_span(const aai_plan &p, int64_t y, int64_t &xa, int64_t &xb) {
    // a pixel is "covered" when floor(c - reach - 1) <= mod-1 and ceil(c + reach + 1) >= 0 on both axes;
    // use the slightly wider closed condition  -(reach+2) <= c <= mod-1+(reach+2)  (exact count is not needed,
    // only a balanced split and an upper bound for the halo).
    const double m = p.reach + 2.0;
    double c0x, c0y;
    centre_of(p, 0.0, (double)y, c0x, c0y);
    const double dxx = p.side * p.cos_t, dxy = -p.side * p.sin_t;  // centre step per canvas x
    double lo = 0.0, hi = (double)p.dst_w;                         // real-valued x range [lo, hi)
    auto clip = [&](double c0, double d, double cmin, double cmax) {
        // cmin <= c0 + x*d <= cmax
        if (std::fabs(d) < 1e-300) {
            if (c0 < cmin || c0 > cmax) hi = lo;
            return;
        }
        double t0 = (cmin - c0) / d, t1 = (cmax - c0) / d;
        if (t0 > t1) {
            double t = t0;
            t0 = t1;
            t1 = t;
        }
        if (t0 > lo) lo = t0;
        if (t1 + 1.0 < hi) hi = t1 + 1.0;
    };
    clip(c0x, dxx, -m, (double)(p.mod_w - 1) + m);
    clip(c0y, dxy, -m, (double)(p.mod_h - 1) + m);
    if (!(hi > lo)) {
        xa = xb = 0;
        return;
    }
    xa = (int64_t)std::floor(lo);
    xb = (int64_t)std::ceil(hi);
    if (xa < 0) xa = 0;
    if (xb > p.dst_w) xb = p.dst_w;
    if (xb < xa) xb = xa;
}

}  // namespace

extern "C" {

const char *aai_status_string(int status) {
    switch (status) {
        case AAI_OK: return "";
        case AAI_ERR_RESOLUTION_XY: return "Assumed X & Y resolution are same.";
        case AAI_ERR_RESOLUTION_NONPOS: return "0 or negative resolution is not acceptable.";
        case AAI_ERR_NO_ROWS: return "There is no data in src array.";
        case AAI_ERR_NO_COLUMNS: return "There is no data in the second dimension of src array.";
        case AAI_ERR_ANGLE: return "Rotation angle is not a finite number.";
        case AAI_ERR_ARGUMENT: return "Invalid argument.";
        case AAI_ERR_CUDA: return "CUDA failure.";
        case AAI_ERR_NO_DEVICE: return "No usable CUDA device (this library has no CPU fallback).";
        default: return "Unknown status.";
    }
}

int aai_plan_create(int64_t src_w, int64_t src_h, double src_res_x, double src_res_y, double dst_res_x,
                    double dst_res_y, double src_iso_x, double src_iso_y, double angle_deg, aai_plan *plan) {
    if (!plan) return AAI_ERR_ARGUMENT;
    std::memset(plan, 0, sizeof *plan);
    aai_plan &p = *plan;
    // the four reference checks, in the reference's order (Source.cpp:112-132)
    if (DBL_EPSILON < std::fabs(src_res_x - src_res_y) || DBL_EPSILON < std::fabs(dst_res_x - dst_res_y))
        return p.status = AAI_ERR_RESOLUTION_XY;
    if (src_res_x <= DBL_EPSILON || dst_res_x <= DBL_EPSILON) return p.status = AAI_ERR_RESOLUTION_NONPOS;
    if (src_h <= 0) return p.status = AAI_ERR_NO_ROWS;
    if (src_w <= 0) return p.status = AAI_ERR_NO_COLUMNS;
    // beyond the reference: it never terminates on a non-finite angle (141-142), NaN resolutions fall
    // through its comparisons
    if (!std::isfinite(angle_deg)) return p.status = AAI_ERR_ANGLE;
    if (!std::isfinite(src_res_x) || !std::isfinite(dst_res_x) || !std::isfinite(src_iso_x) ||
        !std::isfinite(src_iso_y))
        return p.status = AAI_ERR_ARGUMENT;

    // expansion factor and quadrant reduction (139-146)
    const double scale_real = dst_res_x / src_res_x * std::sqrt(2) + 1 + DBL_EPSILON;
    if (!(scale_real < 65536.0)) return p.status = AAI_ERR_ARGUMENT;
    p.scale = static_cast<unsigned int>(scale_real);
    double angle = angle_deg;
    if (std::fabs(angle) > 1e9) angle = std::fmod(angle, 360.0);  // same value the +-360 loops reach, in O(1)
    while (angle < 0) angle += 360;
    while (360 <= angle) angle -= 360;
    if (angle < 90) {
        p.quadrant = 0;
    } else if (angle < 180) {
        p.quadrant = 1;
        angle -= 90;
    } else if (angle < 270) {
        p.quadrant = 2;
        angle -= 180;
    } else {
        p.quadrant = 3;
        angle -= 270;
    }
    p.theta_deg = angle;
    p.sin_t = std::sin(angle / 180.0 * M_PI);
    p.cos_t = std::cos(angle / 180.0 * M_PI);

    p.src_w = src_w;
    p.src_h = src_h;
    const int64_t ew = src_w * (int64_t)p.scale, eh = src_h * (int64_t)p.scale;
    if (ew > 0x7fffffffLL || eh > 0x7fffffffLL) return p.status = AAI_ERR_ARGUMENT;
    const bool swapped = (p.quadrant == 1 || p.quadrant == 3);
    const unsigned mw = (unsigned)(swapped ? eh : ew), mh = (unsigned)(swapped ? ew : eh);  // (150-156)
    p.mod_w = mw;
    p.mod_h = mh;

    // isocentre and resolution in the expanded frame (173-178); the isocentre is NOT quadrant-rotated
    p.iso_x = src_iso_x * p.scale + (p.scale - 1) / 2.0;
    p.iso_y = src_iso_y * p.scale + (p.scale - 1) / 2.0;
    const double src_res = src_res_x * p.scale;
    p.ratio = dst_res_x / src_res;
    p.side = src_res / dst_res_x;
    // canvas size (179-180)
    const double wreal = std::round((mw * std::fabs(p.cos_t) + mh * std::fabs(p.sin_t)) * p.ratio);
    const double hreal = std::round((mw * std::fabs(p.sin_t) + mh * std::fabs(p.cos_t)) * p.ratio);
    if (!(wreal < 2147483647.0) || !(hreal < 2147483647.0)) return p.status = AAI_ERR_ARGUMENT;
    p.dst_w = (unsigned int)wreal;
    p.dst_h = (unsigned int)hreal;
    // returned isocentre (truncated) and its fractional part (181-186)
    const double dix = (p.iso_x * p.cos_t + (mh - p.iso_y) * p.sin_t) * p.ratio;
    const double diy = (p.iso_x * p.sin_t + p.iso_y * p.cos_t) * p.ratio;
    if (!(std::fabs(dix) < 2147483647.0) || !(std::fabs(diy) < 2147483647.0)) return p.status = AAI_ERR_ARGUMENT;
    p.off_ix = dix - int(dix);
    p.off_iy = diy - int(diy);
    p.dst_iso_x = (int)dix;
    p.dst_iso_y = (int)diy;
    // canvas offset = min(0, the four rotated corners) (187-200)
    const double sx = p.iso_x, sy = p.iso_y, c = p.cos_t, s = p.sin_t;
    double ox = 0, oy = 0;
    ox = std::fmin(ox, -sx * c + sy * s + sx);
    oy = std::fmin(oy, -sx * s - sy * c + sy);
    ox = std::fmin(ox, (mw - 1 - sx) * c + sy * s + sx);
    oy = std::fmin(oy, (mw - 1 - sx) * s - sy * c + sy);
    ox = std::fmin(ox, -sx * c - (mh - 1 - sy) * s + sx);
    oy = std::fmin(oy, -sx * s + (mh - 1 - sy) * c + sy);
    ox = std::fmin(ox, (mw - 1 - sx) * c - (mh - 1 - sy) * s + sx);
    oy = std::fmin(oy, (mw - 1 - sx) * s + (mh - 1 - sy) * c + sy);
    p.off_x = ox;
    p.off_y = oy;
    p.reach = p.side * std::sqrt(2) / 2;  // as written at 426-429
    // The reference flattens its edge lines when |tan| < DBL_EPSILON (240); at exactly 0 the footprints are
    // axis-aligned squares and the overlap factorises per axis (separable path).
    p.axis_aligned = (p.sin_t == 0.0 && p.cos_t == 1.0) ? 1 : 0;
    return p.status = AAI_OK;
}

int64_t aai_covered_pixels(const aai_plan *plan, int64_t row0, int64_t row1) {
    if (!plan || plan->status != AAI_OK) return 0;
    if (row0 < 0) row0 = 0;
    if (row1 > plan->dst_h) row1 = plan->dst_h;
    int64_t total = 0;
    for (int64_t y = row0; y < row1; ++y) {
        int64_t xa, xb;
        covered_span(*plan, y, xa, xb);
        total += xb - xa;
    }
    return total;
}

// Cost of an empty canvas pixel relative to a covered one.  An empty pixel still costs its centre / range set-up and its
// zero store (~150 instructions), a covered one 940 (FP32 overlap kernel, BASELINE config 4), 740 (its upscaling path,
// config 3), ~2650 (FP64 kernel) or ~330 (fast mode); rows near the canvas corners also carry more partially covered
// warps and border pixels.  Fitted to per-band kernel times on B200, t = a (covered + w empty) + c per launch, with the
// final kernels of round 2 (tools/dev_bands.py on one GPU; 8-GPU band times of profiles/r2_zz_scale_n8.json):
// w = 0.17 (config 4: 0.166 ... 0.171), 0.20 (config 3), 0.06 (FP64 kernel), 0.27 (fast mode).
double aai_band_empty_weight(const aai_plan *plan, int mode, int arith) {
    if (mode == AAI_MODE_FAST) return 0.27;
    if (arith == AAI_ARITH_F64) return 0.06;
    return (plan && plan->scale >= 3) ? 0.20 : 0.17;
}

int aai_partition_rows(const aai_plan *plan, int n_parts, int64_t *bounds) {
    return aai_partition_rows_weighted(plan, n_parts, aai_band_empty_weight(plan, AAI_MODE_AREA_AVERAGE, AAI_ARITH_F32), bounds);
}

int aai_partition_rows_weighted(const aai_plan *plan, int n_parts, double empty_weight, int64_t *bounds) {
    if (!plan || plan->status != AAI_OK || n_parts <= 0 || !bounds || !(empty_weight >= 0.0) || !(empty_weight <= 1.0))
        return AAI_ERR_ARGUMENT;
    const int64_t h = plan->dst_h;
    // weight of a row = covered pixels + empty_weight per canvas pixel outside the rotated image; the +1 keeps the split
    // defined for fully empty canvases
    std::vector<double> prefix((size_t)h + 1, 0.0);
    for (int64_t y = 0; y < h; ++y) {
        int64_t xa, xb;
        covered_span(*plan, y, xa, xb);
        prefix[(size_t)y + 1] = prefix[(size_t)y] + (double)(xb - xa) + empty_weight * (double)(plan->dst_w - (xb - xa)) + 1.0;
    }
    const double total = prefix[(size_t)h];
    bounds[0] = 0;
    int64_t y = 0;
    for (int k = 1; k < n_parts; ++k) {
        const double target = total * (double)k / (double)n_parts;
        while (y < h && prefix[(size_t)y + 1] <= target) ++y;
        // choose the closer of y / y+1
        int64_t cut = y;
        if (y < h && (target - prefix[(size_t)y]) > (prefix[(size_t)y + 1] - target)) cut = y + 1;
        if (cut < bounds[k - 1]) cut = bounds[k - 1];
        bounds[k] = cut;
    }
    bounds[n_parts] = h;
    return AAI_OK;
}

int aai_band_source_window(const aai_plan *plan, int64_t row0, int64_t row1, int64_t *src_x0, int64_t *src_x1,
                           int64_t *src_y0, int64_t *src_y1) {
    if (!plan || plan->status != AAI_OK) return AAI_ERR_ARGUMENT;
    const aai_plan &p = *plan;
    if (row0 < 0) row0 = 0;
    if (row1 > p.dst_h) row1 = p.dst_h;
    int64_t mx0 = 0, mx1 = 0, my0 = 0, my1 = 0;  // half-open, expanded frame
    if (row1 > row0 && p.dst_w > 0) {
        // the centre map is affine, so the band's centres lie in the parallelogram spanned by its 4 corners
        double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
        const double xs[2] = {0.0, (double)(p.dst_w - 1)}, ys[2] = {(double)row0, (double)(row1 - 1)};
        for (double x : xs)
            for (double y : ys) {
                double cx, cy;
                centre_of(p, x, y, cx, cy);
                lox = std::fmin(lox, cx);
                hix = std::fmax(hix, cx);
                loy = std::fmin(loy, cy);
                hiy = std::fmax(hiy, cy);
            }
        const double m = p.reach + 2.0;  // search window (426-429) plus one pixel of slack
        mx0 = (int64_t)std::floor(std::fmax(lox - m, 0.0));
        my0 = (int64_t)std::floor(std::fmax(loy - m, 0.0));
        mx1 = (int64_t)std::fmin(std::ceil(hix + m) + 1.0, (double)p.mod_w);
        my1 = (int64_t)std::fmin(std::ceil(hiy + m) + 1.0, (double)p.mod_h);
        if (mx1 < mx0) mx1 = mx0;
        if (my1 < my0) my1 = my0;
    }
    // expanded+rotated frame -> expanded frame (inverse of Source.cpp:163-168) -> original pixels
    int64_t ex0, ex1, ey0, ey1;
    switch (p.quadrant) {
        case 0: ex0 = mx0; ex1 = mx1; ey0 = my0; ey1 = my1; break;
        case 1: ex0 = my0; ex1 = my1; ey0 = p.mod_w - mx1; ey1 = p.mod_w - mx0; break;
        case 2: ex0 = p.mod_w - mx1; ex1 = p.mod_w - mx0; ey0 = p.mod_h - my1; ey1 = p.mod_h - my0; break;
        default: ex0 = p.mod_h - my1; ex1 = p.mod_h - my0; ey0 = mx0; ey1 = mx1; break;
    }
    const int64_t S = (int64_t)p.scale;
    auto lo = [&](int64_t e) { return e / S; };
    auto hi = [&](int64_t e) { return (e + S - 1) / S; };
    int64_t x0 = lo(ex0), x1 = hi(ex1), y0 = lo(ey0), y1 = hi(ey1);
    if (x1 > p.src_w) x1 = p.src_w;
    if (y1 > p.src_h) y1 = p.src_h;
    if (x1 <= x0 || y1 <= y0) x0 = x1 = y0 = y1 = 0;
    if (src_x0) *src_x0 = x0;
    if (src_x1) *src_x1 = x1;
    if (src_y0) *src_y0 = y0;
    if (src_y1) *src_y1 = y1;
    return AAI_OK;
}

}  // extern "C"
