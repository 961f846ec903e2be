// TEST-ONLY host build of the device cell arithmetic (csrc/aai_cell.cuh), so that the CPU suite can compare
// the kernels' closed form with the oracle pair by pair without a GPU.  Never linked into the product.
#include <cmath>
#include "../area_average_interpolation_b200/csrc/aai_cell.cuh"

static AaiShape make_shape(double c, double s, double L) {
    AaiShape g;
    g.cs = c; g.sn = s; g.half = L / 2;
    g.hc = g.half * c; g.hs = g.half * s;
    g.k_sc = s / c; g.k_hc = g.half / c; g.k_cs = c / s; g.k_hs = g.half / s;
    g.inv_c = 1.0 / c; g.inv_s = 1.0 / s;
    g.m = (c + s) / 2; g.thr = std::fabs(c - s) / 2;
    return g;
}

extern "C" void aai_test_pair_areas(double c, double s, double L, const double *cx, const double *cy, const int *i,
                                    const int *j, double *out, long long n) {
    const AaiShape g = make_shape(c, s, L);
    for (long long k = 0; k < n; ++k) out[k] = aai_pair_area(g, cx[k], cy[k], i[k], j[k]);
}
