// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Wraps the UNMODIFIED upstream translation unit (Source.cpp) in an extern "C"
// shim so that tests/ and bench.py's cpu_baseline / --impl reference legs can call
// the reference's own public operator.  No upstream text lives in this repo: the
// upstream file is #include-d from where it lies (path given by the build recipe,
// oracle/Makefile, through -DAAI_REFERENCE_SOURCE=...), and the output goes to
// oracle/_ref/ (git-ignored).
//
// Entry points wrapped:
//   AreaAverageInterpolation::areaAverageInterpolation      (Source.cpp:55-583)
//   AreaAverageInterpolation::fastAreaAverageInterpolation  (Source.cpp:584-911)
//   AreaAverageInterpolation::getArea / getIntersectionType (private, 986-1431) via
//   the per-pair probe below (std headers are included first so that the
//   `private -> public` switch only affects the upstream class).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#define private public
#define main aai_reference_main
#include AAI_REFERENCE_SOURCE
#undef main
#undef private

namespace {
struct CoutSilencer {
    std::streambuf *saved;
    std::ios::fmtflags flags;
    std::streamsize prec;
    CoutSilencer() : saved(std::cout.rdbuf(nullptr)), flags(std::cout.flags()), prec(std::cout.precision()) {}
    ~CoutSilencer() {
        std::cout.rdbuf(saved);
        std::cout.clear();
        std::cout.flags(flags);
        std::cout.precision(prec);
    }
};
thread_local IMG g_dst;
}  // namespace

extern "C" {

// mode 1 = areaAverageInterpolation, mode 2 = fastAreaAverageInterpolation.
// src is row-major h x w doubles (w may be 0 / h may be 0 to exercise the error paths).
// Returns 1 on success (reference ret.first), 0 on failure; msg receives ret.second.
// The result is kept in a thread-local buffer; read it back with aai_ref_fetch().
int aai_ref_run(int mode, const double *src, int w, int h, double srcResX, double srcResY, double dstResX,
                double dstResY, double isoX, double isoY, double angleDeg, int *dstW, int *dstH, double *dstIsoX,
                double *dstIsoY, char *msg, int msgCap, double *seconds) {
    IMG s(h > 0 ? h : 0);
    for (int y = 0; y < h; ++y) s[y].assign(src + (size_t)y * w, src + (size_t)y * w + w);
    dP dstIso = std::make_pair(*dstIsoX, *dstIsoY);
    std::pair<bool, std::string> ret;
    AreaAverageInterpolation aa;
    g_dst.clear();
    {
        CoutSilencer quiet;
        auto t0 = std::chrono::steady_clock::now();
        if (mode == 2)
            ret = aa.fastAreaAverageInterpolation(s, g_dst, std::make_pair(srcResX, srcResY),
                                                  std::make_pair(dstResX, dstResY), std::make_pair(isoX, isoY), dstIso,
                                                  angleDeg);
        else
            ret = aa.areaAverageInterpolation(s, g_dst, std::make_pair(srcResX, srcResY),
                                              std::make_pair(dstResX, dstResY), std::make_pair(isoX, isoY), dstIso,
                                              angleDeg);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    *dstIsoX = dstIso.first;
    *dstIsoY = dstIso.second;
    *dstH = (int)g_dst.size();
    *dstW = g_dst.empty() ? 0 : (int)g_dst.front().size();
    if (msg && msgCap > 0) {
        std::strncpy(msg, ret.second.c_str(), (size_t)msgCap - 1);
        msg[msgCap - 1] = 0;
    }
    return ret.first ? 1 : 0;
}

// Copies the last result of this thread (row-major dstH x dstW doubles).
void aai_ref_fetch(double *out) {
    size_t k = 0;
    for (auto &row : g_dst)
        for (double v : row) out[k++] = v;
}

void aai_ref_release() { IMG().swap(g_dst); }

// Segment/segment classifier of the reference (Source.cpp:986-1034), exposed for unit tests.
int aai_ref_intersection_type(double p1x, double p1y, double p2x, double p2y, double q1x, double q1y, double q2x,
                              double q2y, double *r, double *s) {
    AreaAverageInterpolation aa;
    return aa.getIntersectionType(std::make_pair(p1x, p1y), std::make_pair(p2x, p2y), *r, std::make_pair(q1x, q1y),
                                  std::make_pair(q2x, q2y), *s);
}

// The reference's own area dispatcher (Source.cpp:1035-1431) on a caller-built pixel state.
// xa/xb/ya/yb: hit lists (n* entries each), centre / vertex flags, vertex offset from the pixel's top-left corner.
double aai_ref_get_area(const double *xa, int nxa, const double *xb, int nxb, const double *ya, int nya,
                        const double *yb, int nyb, int centreIn, int vertexIn, double vx, double vy) {
    AreaAverageInterpolation aa;
    AreaAverageInterpolation::PixelState st;
    st.intersections["xa"] = std::vector<double>(xa, xa + nxa);
    st.intersections["xb"] = std::vector<double>(xb, xb + nxb);
    st.intersections["ya"] = std::vector<double>(ya, ya + nya);
    st.intersections["yb"] = std::vector<double>(yb, yb + nyb);
    st.xCounts = (unsigned char)(nxa + nxb);
    st.yCounts = (unsigned char)(nya + nyb);
    st.isIncludedSrcPixelCenter = centreIn != 0;
    st.isIncludedDstPixelVertex = vertexIn != 0;
    st.vertexPos = std::make_pair(vx, vy);
    return aa.getArea(st);
}

}  // extern "C"
