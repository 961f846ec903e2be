// Per-(footprint, unit cell) overlap arithmetic shared by every overlap kernel.
//
// Host+device inline functions, so that the CPU test-suite can exercise exactly the arithmetic the kernels
// run (tests/cell_math_host.cpp compiles this header with g++ for unit tests only; the product path never
// evaluates it on the host).
//
// Geometry: the footprint of a canvas pixel is the square |u| <= h, |v| <= h with
//   u = (p - C).(c, -s),  v = (p - C).(s, c),  h = L/2,  theta in (0, 90 deg), y pointing down;
// its "left/right" edges u = +-h have direction (s, c) (the reference's "vertical lines", Source.cpp:457-468),
// its "top/bottom" edges v = +-h have direction (c, -s).  A cell is the unit square centred at integer (i, j).
#ifndef AAI_CELL_CUH_
#define AAI_CELL_CUH_

#include <math.h>

#if defined(__CUDACC__)
#define AAI_HD __host__ __device__ __forceinline__
#else
#define AAI_HD inline
#endif

// image-wide constants of the footprint shape (all FP64, computed once on the host)
struct AaiShape {
    double cs, sn;     // cos, sin of the reduced angle
    double half;       // h
    double hc, hs;     // h*c, h*s
    double k_sc, k_hc; // s/c, h/c
    double k_cs, k_hs; // c/s, h/s
    double inv_c, inv_s;
    double m;          // (c+s)/2: half extent of a unit cell along u or v
    double thr;        // |c-s|/2: a line isolates exactly one cell corner iff thr < |dist| < m
};

// chord of the footprint on the horizontal grid line y = Cy + ty:  x in Cx + [xl, xr]  (empty if xl > xr)
AAI_HD void aai_chord_h(const AaiShape &g, double ty, double &xl, double &xr) {
    const double p = ty * g.k_sc, q = ty * g.k_cs;
    xl = fmax(p - g.k_hc, -q - g.k_hs);
    xr = fmin(p + g.k_hc, g.k_hs - q);
}
// chord of the footprint on the vertical grid line x = Cx + tx:  y in Cy + [yt, yb]
AAI_HD void aai_chord_v(const AaiShape &g, double tx, double &yt, double &yb) {
    const double p = tx * g.k_cs, q = tx * g.k_sc;
    yt = fmax(p - g.k_hs, -q - g.k_hc);
    yb = fmin(p + g.k_hs, g.k_hc - q);
}
// length of [lo, hi] ∩ [r - 1/2, r + 1/2]
AAI_HD double aai_overlap1(double lo, double hi, double r) {
    return fmax(fmin(hi, r + 0.5) - fmax(lo, r - 0.5), 0.0);
}

// Overlap area of the footprint with the unit cell whose centre is (rx, ry) relative to the footprint centre,
// given the lengths of the four cell sides inside the footprint.  Reference-compatible (includes the shape 2/4
// leg quirk of Source.cpp:1055-1062).
AAI_HD double aai_cell_area(const AaiShape &g, double rx, double ry, double lenT, double lenB, double lenL,
                            double lenR) {
    // footprint-local coordinates of the cell centre; nearest footprint vertex V in cell-local coordinates
    const double u0 = rx * g.cs - ry * g.sn;
    const double v0 = rx * g.sn + ry * g.cs;
    const double vx = (copysign(g.hc, u0) + copysign(g.hs, v0)) - rx;
    const double vy = (copysign(g.hc, v0) - copysign(g.hs, u0)) - ry;
    // Green's theorem about V: A = 1/2 sum_sides dist(V, side) * len(side ∩ footprint)
    double area = 0.25 * ((lenT + lenB) + (lenL + lenR)) + 0.5 * (vy * (lenT - lenB) + vx * (lenL - lenR));
    // Reference quirk.  a = signed distance of the cell centre inside the nearest left/right edge.
    const double a = g.half - fabs(u0);
    const double aa = fabs(a);
    if (aa > g.thr && aa < g.m) {  // that edge's line isolates exactly one cell corner
        const double sv = copysign(1.0, v0), su = copysign(1.0, u0);
        // the left/right edge is the ray from W = sv*V along -(s,c); the cell is [-1/2,1/2]^2 (slab test)
        const double wx = sv * vx, wy = sv * vy;
        const double u_in = fmax((wx - 0.5) * g.inv_s, (wy - 0.5) * g.inv_c);
        const double u_out = fmin((wx + 0.5) * g.inv_s, (wy + 0.5) * g.inv_c);
        // the top/bottom edge is the ray from Z = su*V along (-c, s)
        const double zx = su * vx, zy = su * vy;
        const double v_in = fmax((zx - 0.5) * g.inv_c, (-0.5 - zy) * g.inv_s);
        const double v_out = fmin((zx + 0.5) * g.inv_c, (0.5 - zy) * g.inv_s);
        const bool u_through = u_in > 0.0 && u_in < u_out;  // crosses two cell sides inside the edge segment
        const bool v_hits = v_in < v_out && v_out > 0.0;    // top/bottom edge segment meets the cell
        if (u_through && !v_hits) {
            if (a < 0.0) {  // one corner inside, at distance d from the edge: reference shape 2
                const double d = a + g.m;
                area = 0.5 * (1.0 - d * g.inv_c) * (1.0 - d * g.inv_s);
            } else {  // one corner outside (d < 0): reference shape 4
                const double d = a - g.m;
                area = 1.0 - 0.5 * (1.0 + d * g.inv_c) * (1.0 + d * g.inv_s);
            }
        }
    }
    return area;
}

// Stand-alone form for one (footprint centre, cell) pair: computes the four chords itself.
AAI_HD double aai_pair_area(const AaiShape &g, double cx, double cy, int i, int j) {
    const double rx = (double)i - cx, ry = (double)j - cy;
    double xlT, xrT, xlB, xrB, ytL, ybL, ytR, ybR;
    aai_chord_h(g, ry - 0.5, xlT, xrT);
    aai_chord_h(g, ry + 0.5, xlB, xrB);
    aai_chord_v(g, rx - 0.5, ytL, ybL);
    aai_chord_v(g, rx + 0.5, ytR, ybR);
    return aai_cell_area(g, rx, ry, aai_overlap1(xlT, xrT, rx), aai_overlap1(xlB, xrB, rx),
                         aai_overlap1(ytL, ybL, ry), aai_overlap1(ytR, ybR, ry));
}

#endif  // AAI_CELL_CUH_
