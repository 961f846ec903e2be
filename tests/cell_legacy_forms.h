// TEST-ONLY cross-check formulations of the FP32 cell arithmetic (moved out of the product header
// area_average_interpolation_b200/csrc/aai_cell.cuh in round 2: no kernel calls them).  They are earlier formulations of
// the reference's shape-2/4 quirk -- a per-CELL decision form (scalar and packed) and a per-ROW form -- kept as
// independent checks of the per-EDGE form the kernels run (aai_edge_quirk_*), see tests/test_cell_math_host.py.
#ifndef AAI_CELL_LEGACY_TEST_H_
#define AAI_CELL_LEGACY_TEST_H_

#include "../area_average_interpolation_b200/csrc/aai_cell.cuh"

AAI_HD bool aai_sign_product_positive(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return (__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c)) >= 0;
#else
    return ((signbit(a) ? 1 : 0) ^ (signbit(b) ? 1 : 0) ^ (signbit(c) ? 1 : 0)) == 0;
#endif
}

// same, also returning where the left / right edge LINES (u = -h / u = +h) cross the grid line
AAI_HD void aai_chord_h_f32(const AaiShapeF &g, float ty, float &xl, float &xr, float &line_l, float &line_r) {
    line_l = fmaf(ty, g.k_sc, -g.k_hc);
    line_r = fmaf(ty, g.k_sc, g.k_hc);
    xl = fmaxf(line_l, fmaf(-ty, g.k_cs, -g.k_hs));
    xr = fminf(line_r, fmaf(-ty, g.k_cs, g.k_hs));
    xr = fmaxf(xr, xl);
}

// (u0, v0): footprint-local coordinates of the cell centre.  `worst` accumulates the smallest |margin| of the
// quirk decision over the cells of a pixel: the caller redoes the pixel in FP64 when worst < g.tau.
//
// Quirk decision, branch-free (derivation in DESIGN.md §3.3).  Work in the frame W = sv * (cell-local), where the
// nearest left/right edge is the ray from W = sv*V along -(s,c) and the nearest top/bottom edge the ray from W along
// rho*(-c,s), rho = su*sv.  The edge line isolates exactly one corner iff thr < |a| < m; that corner is the
// top-right one (lambda = rho*sign(a) > 0) or the bottom-left one (lambda < 0).  With (p,q,kk) = (wx,wy,s/c) resp.
// (wy,wx,c/s):  both crossings lie on the edge SEGMENT iff p > 1/2 and q > -1/2, and the top/bottom edge misses the
// cell iff a < 0 (it runs away from the cell) or q + kk (p - 1/2) > 1/2.
AAI_HD float aai_cell_area_f32(const AaiShapeF &g, float u0, float v0, float lenT, float lenB, float lenL,
                               float lenR, float &worst) {
    const float ca = copysignf(g.half, u0) - u0;  // V - cell centre, along u
    const float cb = copysignf(g.half, v0) - v0;  // V - cell centre, along v
    const float vx = fmaf(ca, g.cs, cb * g.sn);
    const float vy = fmaf(cb, g.cs, -ca * g.sn);
    const float area = fmaf(0.25f, (lenT + lenB) + (lenL + lenR), 0.5f * fmaf(vy, lenT - lenB, vx * (lenL - lenR)));
    const float a = g.half - fabsf(u0);
    const float aa = fabsf(a);
    const float wx = copysignf(1.0f, v0) * vx, wy = copysignf(1.0f, v0) * vy;
    // lambda = su*sv*sign(a) > 0: the isolated corner is the top-right one (W frame).  Taken from the SIGN BITS so
    // that it stays consistent with copysign() above when u0 or v0 is exactly +-0 (symmetric configurations).
    const bool tr = aai_sign_product_positive(u0, v0, a);
    const float p = tr ? wx : wy, q = tr ? wy : wx, kk = tr ? g.k_sc : g.k_cs;
    const float m2 = g.m - aa;
    const float m3 = p - 0.5f;
    const float m5 = a < 0.0f ? 1.0f : fmaf(kk, m3, q - 0.5f);
    const float need = fminf(fminf(fminf(aa - g.thr, m2), fminf(m3, q + 0.5f)), m5);
    worst = fminf(worst, fabsf(need));
    // reference shape 2 (one corner inside, a < 0) / shape 4 (one corner outside): legs 1 - t/c and 1 - t/s
    const float tri = 0.5f * fmaf(-m2, g.inv_c, 1.0f) * fmaf(-m2, g.inv_s, 1.0f);
    const float quirk = a < 0.0f ? tri : 1.0f - tri;
    return need > 0.0f ? quirk : area;
}

AAI_HD float aai_flip(float v, float sign_of) {  // v with its sign flipped when sign_of is negative (one LOP3)
#if defined(__CUDA_ARCH__)
    return __int_as_float(__float_as_int(v) ^ (__float_as_int(sign_of) & 0x80000000));
#else
    return signbit(sign_of) ? -v : v;
#endif
}

// cells k (lane .x) and k+1 (lane .y) of one row; arguments as in aai_cell_area_f32
AAI_HD AaiF2 aai_cell_area_f32x2(const AaiShapeF &g, AaiF2 u0, AaiF2 v0, AaiF2 lenT, AaiF2 lenB, AaiF2 lenL,
                                 AaiF2 lenR, float &worst) {
    const AaiF2 ca = aai_sub2(aai_f2(copysignf(g.half, u0.x), copysignf(g.half, u0.y)), u0);  // su * a
    const AaiF2 cb = aai_sub2(aai_f2(copysignf(g.half, v0.x), copysignf(g.half, v0.y)), v0);
    const AaiF2 cs2 = aai_f2(g.cs), sn2 = aai_f2(g.sn);
    const AaiF2 vx = aai_fma2(ca, cs2, aai_mul2(cb, sn2));
    const AaiF2 vy = aai_fma2(cb, cs2, aai_mul2(ca, aai_f2(-g.sn)));
    const AaiF2 sum4 = aai_add2(aai_add2(lenT, lenB), aai_add2(lenL, lenR));
    const AaiF2 cross = aai_fma2(vy, aai_sub2(lenT, lenB), aai_mul2(vx, aai_sub2(lenL, lenR)));
    const AaiF2 area = aai_fma2(aai_f2(0.25f), sum4, aai_mul2(aai_f2(0.5f), cross));
    // a = su * ca: |a| = |ca|, sign(a) = sign(ca) ^ sign(u0)
    const AaiF2 aa = aai_f2(fabsf(ca.x), fabsf(ca.y));
    const bool neg_x = aai_sign_product_positive(ca.x, u0.x, -1.0f);  // a < 0 (sign bits)
    const bool neg_y = aai_sign_product_positive(ca.y, u0.y, -1.0f);
    const AaiF2 wx = aai_f2(aai_flip(vx.x, v0.x), aai_flip(vx.y, v0.y));
    const AaiF2 wy = aai_f2(aai_flip(vy.x, v0.x), aai_flip(vy.y, v0.y));
    // lambda = su*sv*sign(a) = sign(ca)*sign(v0) > 0: the isolated corner is the top-right one (W frame)
    const bool tr_x = aai_sign_product_positive(ca.x, v0.x, 1.0f), tr_y = aai_sign_product_positive(ca.y, v0.y, 1.0f);
    const AaiF2 p = aai_f2(tr_x ? wx.x : wy.x, tr_y ? wx.y : wy.y);
    const AaiF2 q = aai_f2(tr_x ? wy.x : wx.x, tr_y ? wy.y : wx.y);
    const AaiF2 kk = aai_f2(tr_x ? g.k_sc : g.k_cs, tr_y ? g.k_sc : g.k_cs);
    const AaiF2 m1 = aai_add2(aa, aai_f2(-g.thr));
    const AaiF2 m2 = aai_sub2(aai_f2(g.m), aa);
    const AaiF2 m3 = aai_add2(p, aai_f2(-0.5f));
    const AaiF2 m4 = aai_add2(q, aai_f2(0.5f));
    const AaiF2 m5v = aai_fma2(kk, m3, aai_add2(q, aai_f2(-0.5f)));
    const float m5x = neg_x ? 1.0f : m5v.x, m5y = neg_y ? 1.0f : m5v.y;
    const float need_x = fminf(fminf(fminf(m1.x, m2.x), fminf(m3.x, m4.x)), m5x);
    const float need_y = fminf(fminf(fminf(m1.y, m2.y), fminf(m3.y, m4.y)), m5y);
    worst = fminf(worst, fminf(fabsf(need_x), fabsf(need_y)));
    const AaiF2 one = aai_f2(1.0f);
    const AaiF2 tri = aai_mul2(aai_mul2(aai_f2(0.5f), aai_fma2(m2, aai_f2(-g.inv_c), one)), aai_fma2(m2, aai_f2(-g.inv_s), one));
    const AaiF2 pent = aai_sub2(one, tri);
    AaiF2 out;
    out.x = need_x > 0.0f ? (neg_x ? tri.x : pent.x) : area.x;
    out.y = need_y > 0.0f ? (neg_y ? tri.y : pent.y) : area.y;
    return out;
}

// ------------------------------------------------------------------------------------------------------------
// Row formulation of the quirk (the kernel's previous formulation, kept as an independent cross-check of the edge
// formulation below in the CPU tests -- no kernel calls it any more): exact areas for every cell (Green form, no decision),
// plus, per row band and per left/right edge line, at most two area CORRECTIONS.
//
// A left/right edge line has direction (s,c) (down-right).  In a row band it enters through the band's top at
// column coordinate z (in cells from the left boundary of cell 0) and leaves at z + s/c.  If that crosses a vertical
// grid line (floor differs), the line cuts the top-right corner of the first cell kT (leg lx = 1 - frac(z) along the
// top side, ly = lx*c/s down the right side) and the bottom-left corner of the last cell kB (lx' = frac(z + s/c),
// ly' = lx'*c/s); cells in between are crossed side to side (trapezoids, exact).  For a corner cut with legs (lx, ly):
//     reference area - exact area = +1/2 (1 - lx - ly)   if the corner is the only one inside (shape 2),
//                                   -1/2 (1 - lx - ly)   if it is the only one outside (shape 4).
// Left edge (footprint to its right): kT is the "inside" case, kB the "outside" case; right edge: the other way round.
// The correction applies iff no other footprint edge meets the cell:
//   inside case : both crossing points lie on the edge SEGMENT (their y within [ylo, yhi] of the edge);
//   outside case: the whole cell lies in the top/bottom slab, |v0| <= h - m.
// (verified pair by pair against the FP64 slab form and the oracle.)
// ------------------------------------------------------------------------------------------------------------

// One edge line in one row band.  LEFT: the left edge (u = -h), else the right edge (u = +h).
//   z        column coordinate of the line at the band's top (cells from the left boundary of column 0)
//   yT       y of the band's top, relative to the footprint centre
//   rx0,vrow v0 of cell k in this row is (rx0 + k)*sin + vrow
// Outputs: (k_in, d_in) correction for the "corner inside" cell, (k_out, d_out) for the "corner outside" cell;
// k = -1 when there is none.  `worst` accumulates the smallest decision margin.
template <bool LEFT>
AAI_HD void aai_row_quirk_f32(const AaiShapeF &g, float z, float yT, float rx0, float vrow, int &k_in, float &d_in,
                              int &k_out, float &d_out, float &worst) {
    const float zB = z + g.k_sc;
    const float kTf = floorf(z), kBf = floorf(zB);
    const float fT = z - kTf, fB = zB - kBf;
    const float lxT = 1.0f - fT;
    const float dT = fmaf(-lxT, g.hk, 0.5f);  // 1/2 (1 - lx - ly) of the top-right cut in cell kT
    const float dB = fmaf(-fB, g.hk, 0.5f);   // ... of the bottom-left cut in cell kB
    const bool cut = kBf > kTf;
    float val_in, val_out;
    if (LEFT) {  // edge spans y in [y_lf, y_bt]; top-right cut at kT is the inside case
        val_in = fminf(yT - g.y_lf, g.y_bt - fmaf(lxT, g.k_cs, yT));
        val_out = g.hm - fabsf(fmaf(rx0 + kBf, g.sn, vrow));
        k_in = (int)kTf;
        k_out = (int)kBf;
        d_in = dT;
        d_out = -dB;
    } else {  // edge spans y in [-y_bt, -y_lf]; bottom-left cut at kB is the inside case
        const float yB = yT + 1.0f;
        val_in = fminf(fmaf(-fB, g.k_cs, yB) + g.y_bt, -g.y_lf - yB);
        val_out = g.hm - fabsf(fmaf(rx0 + kTf, g.sn, vrow));
        k_in = (int)kBf;
        k_out = (int)kTf;
        d_in = dB;
        d_out = -dT;
    }
    // decisions: does the line cross a vertical grid line in this band (corner proximity), are the cuts un-disturbed
    if (fmaxf(val_in, val_out) > -g.tau) {
        const float prox = fminf(fminf(fT, lxT), fminf(fB, 1.0f - fB));
        worst = fminf(worst, fminf(prox, fminf(fabsf(val_in), fabsf(val_out))));
    }
    if (!(cut && val_in > 0.0f)) k_in = -1;
    if (!(cut && val_out > 0.0f)) k_out = -1;
}

// Stand-alone FP32 form for one pair (tests): (fx, fy) = footprint centre minus the nearest integer lattice point,
// (di, dj) = cell index relative to that lattice point.
AAI_HD float aai_pair_area_f32(const AaiShapeF &g, float fx, float fy, int di, int dj, float &worst) {
    const float rx = (float)di - fx, ry = (float)dj - fy;
    float xlT, xrT, xlB, xrB, ytL, ybL, ytR, ybR;
    aai_chord_h_f32(g, ry - 0.5f, xlT, xrT);
    aai_chord_h_f32(g, ry + 0.5f, xlB, xrB);
    aai_chord_v_f32(g, rx - 0.5f, ytL, ybL);
    aai_chord_v_f32(g, rx + 0.5f, ytR, ybR);
    const float u0 = fmaf(rx, g.cs, -ry * g.sn), v0 = fmaf(rx, g.sn, ry * g.cs);
    return aai_cell_area_f32(g, u0, v0, aai_overlap1_f32(xlT, xrT, rx - 0.5f), aai_overlap1_f32(xlB, xrB, rx - 0.5f),
                             aai_overlap1_f32(ytL, ybL, ry - 0.5f), aai_overlap1_f32(ytR, ybR, ry - 0.5f), worst);
}

#endif  // AAI_CELL_LEGACY_TEST_H_
