// Fast mode by BINNING (fastAreaAverageInterpolation, Source.cpp:866-907) -- the source-centric form of the same
// result, FP32 arithmetic, float single-channel images, identity addressing (scale 1, quadrant 0).
//
// The reference visits every canvas pixel and averages the expanded-source pixels whose CENTRE lies in the pixel's
// footprint (closed point-in-square test, 837-864).  The footprints of neighbouring canvas pixels are the cells of a
// rotated square lattice of pitch L: they tile the plane, so every source pixel belongs to exactly one footprint
// (two or four only when its centre lies on a footprint edge -- a measure-zero event handled below).  Gathering from
// the canvas side makes a warp read one 32-byte sector per lane and per load (rotated footprints), which is what bound
// the gather kernel (fast_kernel_f32u: L1 tag stage 77-89 %, 0.59 ms on BASELINE config 4 where HBM allows 0.20 ms).
// Here the work is turned around:
//   * a CTA owns a source tile of BT_W x BT_R pixels (core + a 2-pixel halo); lane = one source column, so a warp's
//     load is 128 consecutive bytes -- every source byte crosses L1 once, fully coalesced;
//   * each source pixel computes the canvas pixel it belongs to from the INVERSE of the centre map (212-219): canvas
//     coordinates (U, V) relative to the tile's accumulator box, binned with a round-to-nearest (magic-number add);
//   * a lane walks DOWN its column, so consecutive pixels with the same bin are summed in registers and a run is written
//     to shared memory when the bin changes.  A footprint is convex: it meets a source column in ONE run, and it meets at
//     most NS consecutive columns, so the accumulator slot (canvas pixel, column mod NS) is written exactly once per
//     tile -- plain stores, no atomics, and a summation order that depends on nothing but the geometry (bitwise
//     reproducible whatever the band partition or device count);
//   * after the walk the CTA emits the canvas pixels it OWNS (nearest lattice point of the footprint centre inside the
//     tile core; the halo makes their sums complete): slots added in fixed order, divided by the count, stored.
// The bin decision is discontinuous, so it never rests on FP32 rounding: a pixel whose FP32 coordinate comes within the
// guard band of a bin boundary is re-binned from the FP64 inverse map, and if even that is within 1e-9 of the boundary
// (the closed test counts such a pixel in both footprints) the canvas pixels either side are flagged and evaluated by
// the reference's own expression (pixel_fast_f64).  Canvas pixels whose candidate box is clipped by the image border,
// and the empty ones beyond it, are written by fast_border_kernel (same launch sequence, disjoint pixels).
//
// Selected with AAI_ARITH_F32_BINNED, NOT the default: measured on B200 (profiles/r2_w_*) it reads HBM exactly once and
// takes the load off L1 (l1tex 42 % against 89 %), but at ~27 instructions per SOURCE pixel plus the emit phase it
// executes as many instructions as the gather kernel at a lower issue rate -- 0.66 ms against 0.54 ms on BASELINE config 4,
// slower on every ratio / angle tried.
#include <cuda.h>

#include <cmath>

#include "aai_device.cuh"

using namespace aai_dev;

namespace {

constexpr int BT_THREADS = 128;          // 4 warps
constexpr int BT_W = BT_THREADS;         // source columns per tile: one per lane
constexpr int BT_HX = 2;                 // halo columns either side: the candidate box lies within 2 of the nearest lattice point (hb < 2.5)
constexpr int BT_CW = BT_W - 2 * BT_HX;  // core columns: the canvas pixels whose nearest lattice point lies there are OURS
#ifndef AAI_BIN_ROWS
#define AAI_BIN_ROWS 64
#endif
#ifndef AAI_BIN_MIN_CTAS
#define AAI_BIN_MIN_CTAS 5  // resident CTAs per SM the register allocation allows (6: 80 registers with spills, measured slower)
#endif
constexpr int BT_R = AAI_BIN_ROWS;  // source rows per tile (walked by every lane)
constexpr int BT_HY = 2;
constexpr int BT_CR = BT_R - 2 * BT_HY;
constexpr int BT_NG = 4, BT_NB = 4;  // rows per group, groups per block: a buffer is reloaded right after its group is
                                     // binned and used again (BT_NB - 1) groups later (12 rows of work hide the load)
constexpr float BIN_MAGIC = 12582912.0f;     // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer, held in the low mantissa bits
constexpr unsigned BIN_MAGIC_BITS = 0x4B400000u;

struct BinParams {
    double iu0, iui, iuj;  // canvas x coordinate of the expanded-frame point (i, j): U = iu0 + i iui + j iuj (pixel X covers |U - X| <= 1/2)
    double iv0, ivi, ivj;  // canvas y coordinate V likewise
    int tx0, ty0;          // tile indices of blockIdx (0, 0)
    int pitch, height;     // accumulator box of one tile, in canvas pixels
    float lim;             // 1/2 - guard band: an FP32 offset from the bin centre beyond it is decided in FP64
    float bu, bv;          // (float)iuj, (float)ivj: increments of the tile-local coordinates per source row
    float axx, axy, ayx, ayy;  // the centre map's linear part in FP32 (ownership test on tile-local centres)
    int emit_w;                // box columns per emit pass: 64 or 128 (>= pitch)
    // shear of the accumulator box (see TileGeom): kappa = 0 -> none
    float kappa;
    double kappa_d;
    double wmin_di, wmin_dj;   // corner of the loaded region with the smallest V - kappa U
    int box_rows;              // canvas rows the region can reach (= height without the shear)
    double umin_di, umin_dj, vmin_di, vmin_dj;  // corner of the loaded region with the smallest U resp. V (0 or BT_W / BT_R)
};

template <typename T>
__device__ __forceinline__ void store_m(void *row, int idx, float v);
template <>
__device__ __forceinline__ void store_m<float>(void *row, int idx, float v) {
    ((float *)row)[idx] = v;
}
template <>
__device__ __forceinline__ void store_m<uint8_t>(void *row, int idx, float v) {
    ((uint8_t *)row)[idx] = (uint8_t)__float2int_rd(fminf(fmaxf(v + 0.5f, 0.0f), 255.0f));  // round half up, saturate
}

// The candidate box of canvas pixel (x, y): lattice points within hb (+ guard) of the footprint centre, FP32 on the split
// centre exactly as in fast_kernel_f32u.  "interior": the whole box lies inside the image and is at most NS columns wide
// -- the pixels the binning kernel emits; every other pixel belongs to fast_border_kernel.  Both kernels call THIS
// function, so the two sets are complementary by construction.
struct CandBox {
    int irx, iry;  // lattice point nearest to the centre
    bool interior, empty;
};
__device__ __forceinline__ CandBox cand_box(const AaiKernelParams &kp, int ns, int x, int y) {
    const double cx = fma((double)x, kp.aff_xx, fma((double)y, kp.aff_xy, kp.aff_x0));
    const double cy = fma((double)x, kp.aff_yx, fma((double)y, kp.aff_yy, kp.aff_y0));
    CandBox b;
    b.irx = __double2int_rn(cx);
    b.iry = __double2int_rn(cy);
    const float fx = (float)(cx - (double)b.irx), fy = (float)(cy - (double)b.iry);
    const float ext = kp.shapef.hb + 4e-6f;
    const int bx0 = b.irx + __float2int_ru(fx - ext), bx1 = b.irx + __float2int_rd(fx + ext);
    const int by0 = b.iry + __float2int_ru(fy - ext), by1 = b.iry + __float2int_rd(fy + ext);
    b.empty = bx1 < 0 || by1 < 0 || bx0 > kp.mod_w - 1 || by0 > kp.mod_h - 1;
    b.interior = bx0 >= 0 && by0 >= 0 && bx1 <= kp.mod_w - 1 && by1 <= kp.mod_h - 1 && bx1 - bx0 < ns && by1 - by0 < ns;
    return b;
}

// The accumulator box of a tile.  The loaded source region maps to a ROTATED rectangle of the canvas, which fills little
// more than half of its bounding box, and shared memory per tile is what limits the resident warps.  So the box is
// sheared: entry (Xl, Ys) with Ys = Yl - rint(kappa Xl + d) + const, kappa = the slope of the rectangle's long sides --
// any deterministic integer function of Xl would do, as long as the walk and the emit phase evaluate the SAME one
// (shear_rows below, identical instruction sequence) and the box is tall enough for its range (+ slack, host side).
struct TileGeom {
    int X0, Y0;   // canvas pixel of box column 0 / of local row 0 (before the shear)
    float skd;    // frac(kappa X0) - 1/2
    int skc;      // Ys = Yl - shear_rows(Xl) + skc
};
__device__ __forceinline__ float shear_rows(float xf, float kappa, float skd) {
    return (fmaf(xf, kappa, skd) + BIN_MAGIC) - BIN_MAGIC;  // an integer-valued float
}

// FP64 decision for a source pixel whose FP32 coordinate fell into the guard band (rare: ~4e-4 of the pixels)
__device__ __noinline__ float bin_exact(const BinParams &bp, const TileGeom &tg, int i, int j, unsigned char *flags,
                                        int *any_flag) {
    const double U = fma((double)i, bp.iui, fma((double)j, bp.iuj, bp.iu0));
    const double V = fma((double)i, bp.ivi, fma((double)j, bp.ivj, bp.iv0));
    const double ru = rint(U), rv = rint(V);
    const int xl = (int)ru - tg.X0, yl = (int)rv - tg.Y0;
    auto entry = [&](int fxl, int fyl) -> int {  // box index of local pixel (fxl, fyl), or the spare entry
        const int ys = fyl - (int)shear_rows((float)fxl, bp.kappa, tg.skd) + tg.skc;
        return ((unsigned)fxl < (unsigned)bp.pitch && (unsigned)ys < (unsigned)bp.height) ? ys * bp.pitch + fxl
                                                                                          : bp.pitch * bp.height;
    };
    const bool au = fabs(fabs(U - ru) - 0.5) < 1e-9, av = fabs(fabs(V - rv) - 0.5) < 1e-9;
    if (au || av) {  // on a footprint edge as far as FP64 can tell: the reference's own expression decides (both sides)
        const int sx = U >= ru ? 1 : -1, sy = V >= rv ? 1 : -1;
        flags[entry(xl, yl)] = 1;
        if (au) flags[entry(xl + sx, yl)] = 1;
        if (av) flags[entry(xl, yl + sy)] = 1;
        if (au && av) flags[entry(xl + sx, yl + sy)] = 1;
        *any_flag = 1;
    }
    return BIN_MAGIC + (float)entry(xl, yl);  // the key of the main loop: MAGIC + box index
}

// One lane's walk down its source column: the run of pixels that share a bin is summed in registers and written to the
// slot (bin, column mod NS) when the bin changes.  Keys are MAGIC + box index; index nent is a spare entry that takes
// the write of the (empty) run before the first pixel and of anything the FP64 path finds outside the box.  The stores
// are predicated, not branched around: in almost every row some lane of the warp ends a run.
__device__ __forceinline__ uint32_t bin_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NS>
struct Walk {
    float cur, sum;  // key of the current run, its sum
    int cnt;         // and its length
    // shared-space addresses of this lane's slot in entry 0, minus MAGIC_BITS * stride: address = bits(key) * stride + base
    uint32_t sum_base, cnt_base;
#ifdef AAI_BIN_CHECK  // developer build (tools/build_variant.sh bincheck -DAAI_BIN_CHECK=1): trap on a key outside the box
    uint32_t nent_chk;
    __device__ __forceinline__ void check(uint32_t b) const {
        if (b - BIN_MAGIC_BITS > nent_chk) __trap();
    }
#else
    __device__ __forceinline__ void check(uint32_t) const {}
#endif
    __device__ __forceinline__ void init(const float *sums, const unsigned char *cnts, int slot, int nent) {
        cur = BIN_MAGIC + (float)nent;  // the spare entry
        sum = 0.0f;
        cnt = 0;
        sum_base = bin_smem_u32(sums + slot) - BIN_MAGIC_BITS * (uint32_t)(NS * 4);
        cnt_base = bin_smem_u32(cnts + slot) - BIN_MAGIC_BITS * (uint32_t)NS;
#ifdef AAI_BIN_CHECK
        nent_chk = (uint32_t)nent;
#endif
    }
    __device__ __forceinline__ void flush() {
        const uint32_t b = __float_as_uint(cur);
        check(b);
        asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.u8 [%2], %3;" ::"r"(sum_base + b * (NS * 4)), "f"(sum),
                     "r"(cnt_base + b * NS), "r"(cnt));
    }
    __device__ __forceinline__ void add(float key, float v) {
        const uint32_t b = __float_as_uint(cur);
        check(b);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %0, %1;\n\t@p st.shared.f32 [%2], %3;\n\t@p st.shared.u8 [%4], %5;\n\t}" ::"f"(key),
            "f"(cur), "r"(sum_base + b * (NS * 4)), "f"(sum), "r"(cnt_base + b * NS), "r"(cnt));
        const bool chg = key != cur;
        sum = chg ? v : sum + v;
        cnt = chg ? 1 : cnt + 1;
        cur = key;
    }
};

// bin key of the source pixel in row fj (tile-local, as a float) of this lane's column; d = largest FP32 offset from the
// bin centre over the two axes (the guard-band test).  c2.y includes the shear constant skc.
template <bool SK>
__device__ __forceinline__ float bin_key(float fj, float2 b2, float2 c2, float pitchf, float kappa, float skd, float &d) {
    const float2 t = __ffma2_rn(make_float2(fj, fj), b2, c2);
    const float2 m = __fadd2_rn(t, make_float2(BIN_MAGIC, BIN_MAGIC));
    const float2 r = __fadd2_rn(m, make_float2(-BIN_MAGIC, -BIN_MAGIC));
    const float2 o = __ffma2_rn(r, make_float2(-1.0f, -1.0f), t);
    d = fmaxf(fabsf(o.x), fabsf(o.y));
    const float key = fmaf(r.y, pitchf, m.x);  // MAGIC + box index before the shear, exact
    return SK ? fmaf(shear_rows(r.x, kappa, skd), -pitchf, key) : key;
}

template <typename TO, int NS, bool SK>
__global__ void __launch_bounds__(BT_THREADS, AAI_BIN_MIN_CTAS)
    fast_bin_kernel(const __grid_constant__ AaiKernelParams kp, const __grid_constant__ BinParams bp) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    __shared__ int any_flag;
    __shared__ float rcp_tab[64];  // 1 / n, correctly rounded (n = pixels per footprint; 0 -> 0)
    const int nent = bp.pitch * bp.height;
    float *sums = reinterpret_cast<float *>(bin_smem);            // [nent + 1][NS] run sums, slot = source column mod NS
    unsigned char *cnts = bin_smem + (size_t)(nent + 1) * NS * 4;  // [nent + 1][NS] run lengths
    unsigned char *flags = cnts + (size_t)(nent + 1) * NS;         // [nent + 1] 1: evaluate with the reference's FP64 expression
    const int tid = threadIdx.x;
    const int tbx = bp.tx0 + (int)blockIdx.x, tby = bp.ty0 + (int)blockIdx.y;
    const int i0 = tbx * BT_CW - BT_HX, j0 = tby * BT_CR - BT_HY;  // first loaded column / row
    // accumulator box: canvas coordinates of the loaded region (affine map: extremes at its corners; the corner with the
    // smallest U, V resp. V - kappa U is known from the signs of the map: bp.umin_*, bp.vmin_*, bp.wmin_*)
    const double ulo = fma((double)i0 - 0.5 + bp.umin_di, bp.iui, fma((double)j0 - 0.5 + bp.umin_dj, bp.iuj, bp.iu0));
    const double vlo = fma((double)i0 - 0.5 + bp.vmin_di, bp.ivi, fma((double)j0 - 0.5 + bp.vmin_dj, bp.ivj, bp.iv0));
    TileGeom tg;
    tg.X0 = __double2int_rd(ulo) - 1;
    tg.Y0 = __double2int_rd(vlo) - 1;
    tg.skd = 0.0f;
    tg.skc = 0;
    int ylo_c = tg.Y0, yhi_c = tg.Y0 + bp.height;  // canvas rows the box can hold
    if (SK) {
        const double wi = (double)i0 - 0.5 + bp.wmin_di, wj = (double)j0 - 0.5 + bp.wmin_dj;
        const double wlo = fma(wi, bp.ivi, fma(wj, bp.ivj, bp.iv0)) - bp.kappa_d * fma(wi, bp.iui, fma(wj, bp.iuj, bp.iu0));
        const double kx = bp.kappa_d * (double)tg.X0, fk = floor(kx);
        tg.skd = (float)(kx - fk - 0.5);
        tg.skc = tg.Y0 - __double2int_rd(wlo - 0.5 * bp.kappa_d - 1.5) - (int)fk;
        ylo_c = tg.Y0;  // (unsheared bounds: rows [Y0, Y0 + bp.box_rows) are the only ones the region can reach)
        yhi_c = tg.Y0 + bp.box_rows;
    }
    const int X0 = tg.X0, Y0 = tg.Y0;
    if (X0 + bp.pitch <= 0 || X0 >= kp.dst_w || yhi_c <= kp.row0 || ylo_c >= kp.row1) return;  // nothing to emit

    // ---- this lane's source column: the first rows are requested before anything else ------------------------------
    const int i = i0 + tid;
    const int jlo = max(j0, max(0, kp.src_y0)), jhi = min(j0 + BT_R, min(kp.mod_h, kp.src_y0 + kp.src_rows));
    const bool walk = i >= 0 && i < kp.mod_w && jlo < jhi;
    const int nsuper = walk ? (jhi - jlo) / (BT_NB * BT_NG) : 0;  // blocks of BT_NB groups of BT_NG rows
    const int64_t pitch = kp.src_pitch;
    const char *colp = (const char *)kp.src + (int64_t)(jlo - src_row0(kp)) * pitch + (int64_t)i * 4;
    float buf[BT_NB][BT_NG];
    if (nsuper > 0) {  // (one running pointer: the rows are requested in order)
#pragma unroll
        for (int k = 0; k < BT_NB; ++k)
#pragma unroll
            for (int r = 0; r < BT_NG; ++r) {
                buf[k][r] = __ldg((const float *)colp);
                colp += pitch;
            }
    }

    {  // clear the accumulators
        const int n16 = ((nent + 1) * (NS * 5 + 1) + 15) >> 4;
        uint4 *p = reinterpret_cast<uint4 *>(bin_smem);
        for (int k = tid; k < n16; k += BT_THREADS) p[k] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) any_flag = 0;
        if (tid < 64) rcp_tab[tid] = tid ? 1.0f / (float)tid : 0.0f;
    }
    __syncthreads();

    // ---- walk down the column ------------------------------------------------------------------------------------
    const float pitchf = (float)bp.pitch, kappa = bp.kappa;
    if (walk) {
        // tile-local canvas coordinates of this column's pixel in row j0 (FP64, one rounding), increments per row
        const float2 c2 = make_float2(
            (float)(fma((double)i, bp.iui, fma((double)j0, bp.iuj, bp.iu0)) - (double)X0),
            (float)(fma((double)i, bp.ivi, fma((double)j0, bp.ivj, bp.iv0)) - (double)(Y0 - tg.skc)));
        const float2 b2 = make_float2(bp.bu, bp.bv);
        const float lim = bp.lim, skd = tg.skd;
        Walk<NS> w;
        w.init(sums, cnts, i & (NS - 1), nent);
        float fj = (float)(jlo - j0);
        int j = jlo;
        for (int sg = 0; sg < nsuper; ++sg) {
            const bool more = sg + 1 < nsuper;
#pragma unroll
            for (int k = 0; k < BT_NB; ++k) {
                float v[BT_NG], key[BT_NG];
#pragma unroll
                for (int r = 0; r < BT_NG; ++r) v[r] = buf[k][r];
                if (more) {  // this buffer's rows of the next block: in flight while the other buffers are binned
#pragma unroll
                    for (int r = 0; r < BT_NG; ++r) {
                        buf[k][r] = __ldg((const float *)colp);
                        colp += pitch;
                    }
                }
                float dm = 0.0f;
#pragma unroll
                for (int r = 0; r < BT_NG; ++r) {
                    float d;
                    key[r] = bin_key<SK>(fj + (float)(k * BT_NG + r), b2, c2, pitchf, kappa, skd, d);
                    dm = fmaxf(dm, d);
                }
                if (dm > lim) {  // some pixel of the group sits in the guard band: FP64 decides the group (rare)
#pragma unroll
                    for (int r = 0; r < BT_NG; ++r) key[r] = bin_exact(bp, tg, i, j + k * BT_NG + r, flags, &any_flag);
                }
#pragma unroll
                for (int r = 0; r < BT_NG; ++r) w.add(key[r], v[r]);
            }
            fj += (float)(BT_NB * BT_NG);
            j += BT_NB * BT_NG;
        }
        for (; j < jhi; ++j) {  // rows of an incomplete block (image / band edge); colp already points at row j
            float d;
            float key = bin_key<SK>(fj, b2, c2, pitchf, kappa, skd, d);
            if (d > lim) key = bin_exact(bp, tg, i, j, flags, &any_flag);
            w.add(key, __ldg((const float *)colp));
            colp += pitch;
            fj += 1.0f;
        }
        w.flush();
    }
    __syncthreads();

    // ---- emit the canvas pixels this tile owns ---------------------------------------------------------------------
    // A thread keeps one box column xl and steps through the box rows, so that the shear offset, the canvas column and
    // most of the addresses are per-thread constants.  Ownership: the nearest lattice point of the footprint centre lies
    // in the core -- decided on the tile-local FP32 centre, by cand_box() (FP64) within 1e-3 of a core boundary and for
    // tiles on the image border (elsewhere core + halo inside the image means every owned pixel is interior).
    const int pw = bp.emit_w, xl = tid & (pw - 1), rp = BT_THREADS / pw;  // box columns per pass (64 or 128), rows per pass
    const int X = X0 + xl;
    if (xl >= bp.pitch || (unsigned)X >= (unsigned)kp.dst_w) return;
    const int ci0 = i0 + BT_HX, cj0 = j0 + BT_HY;
    const bool inner = i0 >= 0 && i0 + BT_W <= kp.mod_w && j0 >= 0 && j0 + BT_R <= kp.mod_h;
    const bool anyf = any_flag != 0;
    constexpr float G = 1e-3f;
    constexpr float MX = 0.5f * (float)BT_CW - 0.5f, HX = 0.5f * (float)BT_CW;  // nearest lattice point in [0, BT_CW)
    constexpr float MY = 0.5f * (float)BT_CR - 0.5f, HY = 0.5f * (float)BT_CR;  //   <=> |centre - M| < H
    // local canvas row of box row ys in this column: yl = ys + yoff
    const int yoff = SK ? (int)shear_rows((float)xl, kappa, tg.skd) - tg.skc : 0;
    const int ys0 = tid / pw;
    // footprint centre of (xl, yl) relative to the core origin: (cx0 + yl axy, cy0 + yl ayy)
    const float cx0 = fmaf((float)xl, bp.axx, (float)(fma((double)X0, kp.aff_xx, fma((double)Y0, kp.aff_xy, kp.aff_x0)) - (double)ci0));
    const float cy0 = fmaf((float)xl, bp.ayx, (float)(fma((double)X0, kp.aff_yx, fma((double)Y0, kp.aff_yy, kp.aff_y0)) - (double)cj0));
    const uint32_t rcp_s = bin_smem_u32(rcp_tab);
    uint32_t e = (uint32_t)(ys0 * bp.pitch + xl);
    const uint32_t estep = (uint32_t)(rp * bp.pitch);
    int Y = Y0 + ys0 + yoff;
    char *dp = (char *)kp.dst + (int64_t)(Y - dst_row0(kp)) * kp.dst_pitch + (int64_t)X * (int64_t)sizeof(TO);
    const int64_t dstep = (int64_t)rp * kp.dst_pitch;
    float fyl = (float)(ys0 + yoff);
    const float fstep = (float)rp;
    for (int ys = ys0; ys < bp.height; ys += rp, e += estep, Y += rp, dp += dstep, fyl += fstep) {
        const float ox = fabsf(fmaf(fyl, bp.axy, cx0) - MX), oy = fabsf(fmaf(fyl, bp.ayy, cy0) - MY);
        if (!(inner && ox <= HX - G && oy <= HY - G)) {  // not plainly ours
            if (ox > HX + G || oy > HY + G) continue;
            const CandBox b = cand_box(kp, NS, X, Y);
            if (!b.interior || (unsigned)(b.irx - ci0) >= (unsigned)BT_CW || (unsigned)(b.iry - cj0) >= (unsigned)BT_CR)
                continue;
        }
        if ((unsigned)(Y - kp.row0) >= (unsigned)(kp.row1 - kp.row0)) continue;
        if (anyf && flags[e]) {
            int c64;
            double a64[1];
            pixel_fast_f64<float, 1>(kp, X, Y, c64, a64);
            store_m<TO>(dp, 0, c64 > 0 ? (float)(a64[0] / (double)c64) : 0.0f);
        } else {
            float tot = 0.0f;
            unsigned n = 0;
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) {  // slots in fixed order
                const float4 sv = *reinterpret_cast<const float4 *>(sums + (size_t)e * NS + 4 * q);
                const unsigned c4 = *reinterpret_cast<const unsigned *>(cnts + (size_t)e * NS + 4 * q);
                tot = q == 0 ? ((sv.x + sv.y) + sv.z) + sv.w : (((tot + sv.x) + sv.y) + sv.z) + sv.w;
                n = __dp4a(c4, 0x01010101u, n);
            }
            float rn;  // 1 / n from the table (a footprint holds at most nf^2 <= 25 pixels)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rn) : "r"(rcp_s + (n & 63u) * 4u));
            store_m<TO>(dp, 0, tot * rn);
        }
    }
}

// Canvas pixels the binning kernel does not emit: candidate box clipped by the image border (FP64, the reference's own
// expression) or missing the image altogether (0).  A CTA takes a quarter of a group of 8 canvas rows.  The centre map is
// affine, so for a group of rows "every centre at least m inside the image" and "some centre less than m outside it" are
// intervals of x: tiles inside the first are the binning kernel's (skipped), tiles outside the second are zero-filled
// without looking at their pixels, the few in between evaluate cand_box() per pixel.
constexpr int BD_PARTS = 4;  // CTAs per group of 8 canvas rows
// x-interval on which lo <= base(y) + k x <= hi: for ALL rows y in {y0, y1} (inner = true; exact for the rows between,
// the region is convex) or a superset of the x for which it holds on SOME row between y0 and y1 (inner = false)
__device__ __forceinline__ void row_group_interval(const AaiKernelParams &kp, double y0, double y1, double m, bool inner,
                                                   double &xa, double &xb) {
    xa = -1e300;
    xb = 1e300;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double k = a ? kp.aff_yx : kp.aff_xx;
        const double b0 = a ? fma(y0, kp.aff_yy, kp.aff_y0) : fma(y0, kp.aff_xy, kp.aff_x0);
        const double b1 = a ? fma(y1, kp.aff_yy, kp.aff_y0) : fma(y1, kp.aff_xy, kp.aff_x0);
        const double lo = inner ? m : -m, hi = (double)((a ? kp.mod_h : kp.mod_w) - 1) + (inner ? -m : m);
        if (k != 0.0) {
            const double r = 1.0 / k;
            // the two bounds of each row, ordered
            const double p0 = (lo - b0) * r, q0 = (hi - b0) * r, p1 = (lo - b1) * r, q1 = (hi - b1) * r;
            const double l0 = fmin(p0, q0), u0 = fmax(p0, q0), l1 = fmin(p1, q1), u1 = fmax(p1, q1);
            xa = fmax(xa, inner ? fmax(l0, l1) : fmin(l0, l1));
            xb = fmin(xb, inner ? fmin(u0, u1) : fmax(u0, u1));
        } else if (inner ? (b0 < lo || b0 > hi || b1 < lo || b1 > hi) : (fmax(b0, b1) < lo || fmin(b0, b1) > hi)) {
            xb = -1e300;
        }
    }
}
__device__ __forceinline__ int clamp_int(double v, bool up) {
    return v > 2e9 ? 0x7fffffff : (v < -2e9 ? -0x7fffffff : (up ? __double2int_ru(v) : __double2int_rd(v)));
}
template <typename TO>
__global__ void __launch_bounds__(TILE_W *TILE_H)
    fast_border_kernel(const __grid_constant__ AaiKernelParams kp, const int ns) {
    const int yg = kp.row0 + (int)blockIdx.y * TILE_H;
    const int y = yg + (int)threadIdx.y;
    const int ylast = min(yg + TILE_H - 1, kp.row1 - 1);
    double xa, xb, oa, ob;
    row_group_interval(kp, (double)yg, (double)ylast, kp.hb + 2.0, true, xa, xb);    // surely interior
    row_group_interval(kp, (double)yg, (double)ylast, kp.hb + 2.0, false, oa, ob);   // possibly not empty
    const int ixa = clamp_int(xa, true), ixb = clamp_int(xb, false);
    const int ioa = clamp_int(oa, false), iob = clamp_int(ob, true);
    // this CTA's share of the row group, in tiles of TILE_W columns; the interior tiles [ta, tb] are skipped
    const int ntiles = (kp.dst_w + TILE_W - 1) / TILE_W, per = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t0 = (int)blockIdx.x * per, t1 = min(t0 + per, ntiles);
    int ta = ntiles, tb = -1;
    if (ixa <= ixb) {
        ta = ixa <= 0 ? 0 : (ixa + TILE_W - 1) / TILE_W;
        tb = ixb < 0 ? -1 : min((ixb + 1) / TILE_W - 1, ntiles - 1);
        if (ta > tb) {
            ta = ntiles;
            tb = -1;
        }
    }
    if (y >= kp.row1) return;
    char *drow = (char *)kp.dst + (int64_t)(y - dst_row0(kp)) * kp.dst_pitch;
    for (int part = 0; part < 2; ++part) {
        const int pa = part ? max(t0, tb + 1) : t0, pb = part ? t1 : (tb >= ta ? min(t1, ta) : t1);
        if (part && tb < ta) break;  // no interior tiles: the first part covered everything
        for (int t = pa; t < pb; ++t) {
            const int xt = t * TILE_W, x = xt + (int)threadIdx.x;
            if (x >= kp.dst_w) continue;
            if (xt + TILE_W - 1 < ioa || xt > iob) {  // beyond the image by a margin: empty
                store_m<TO>(drow, x, 0.0f);
                continue;
            }
            const CandBox b = cand_box(kp, ns, x, y);
            if (b.interior) continue;
            if (b.empty) {
                store_m<TO>(drow, x, 0.0f);
            } else {
                int c64;
                double a64[1];
                pixel_fast_f64<float, 1>(kp, x, y, c64, a64);
                store_m<TO>(drow, x, c64 > 0 ? (float)(a64[0] / (double)c64) : 0.0f);
            }
        }
    }
}

template <typename TO, int NS, bool SK>
cudaError_t launch_bin(const AaiKernelParams &kp, const BinParams &bp, dim3 grid, size_t smem, cudaStream_t stream) {
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        const cudaError_t e =
            cudaFuncSetAttribute(fast_bin_kernel<TO, NS, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        attr_dev = dev;
    }
    fast_bin_kernel<TO, NS, SK><<<grid, BT_THREADS, smem, stream>>>(kp, bp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int rows = kp.row1 - kp.row0;
    const dim3 bgrid(BD_PARTS, (rows + TILE_H - 1) / TILE_H, grid.z);
    fast_border_kernel<TO><<<bgrid, dim3(TILE_W, TILE_H), 0, stream>>>(kp, NS);
    aai_count_extra_launches(1);  // two kernels per call (the glue counts one)
    return cudaGetLastError();
}
template <typename TO>
cudaError_t launch_bin_to(const AaiKernelParams &kp, const BinParams &bp, int ns, dim3 grid, size_t smem, cudaStream_t st) {
    const bool sk = bp.kappa != 0.0f;
    if (ns == 4) return sk ? launch_bin<TO, 4, true>(kp, bp, grid, smem, st) : launch_bin<TO, 4, false>(kp, bp, grid, smem, st);
    return sk ? launch_bin<TO, 8, true>(kp, bp, grid, smem, st) : launch_bin<TO, 8, false>(kp, bp, grid, smem, st);
}

}  // namespace

// Returns cudaErrorNotSupported when the binning form does not apply (the caller then runs the gather kernel).
int aai_launch_fast_bin(const AaiKernelParams &kp, int src_dtype, int dst_dtype, void *stream) {
    const int rows = kp.row1 - kp.row0;
    if (rows <= 0 || kp.dst_w <= 0) return (int)cudaSuccess;
    if (src_dtype != AAI_F32 || (dst_dtype != AAI_F32 && dst_dtype != AAI_U8) || kp.channels != 1 || kp.scale != 1 ||
        kp.quadrant != 0 || !kp.f32_ok || kp.mod_w >= (1 << 22) || kp.mod_h >= (1 << 22) ||
        (kp.src_pitch & 3) != 0 || (reinterpret_cast<uintptr_t>(kp.src) & 3) != 0)
        return (int)cudaErrorNotSupported;
    const int nf = (int)floor(2.0 * ((double)kp.shapef.hb + 4e-6) + 1e-6) + 1;
    // halo: the candidate box starts at ceil(-1/2 - hb - guard) >= -BT_HX columns from the nearest lattice point
    if (nf > 8 || kp.hb + 0.5 + 1e-4 >= (double)(BT_HX + 1)) return (int)cudaErrorNotSupported;
    BinParams bp;
    const double det = kp.aff_xx * kp.aff_yy - kp.aff_xy * kp.aff_yx;
    if (!(det > 0.0)) return (int)cudaErrorNotSupported;
    bp.iui = kp.aff_yy / det;
    bp.iuj = -kp.aff_xy / det;
    bp.iu0 = -(kp.aff_yy * kp.aff_x0 - kp.aff_xy * kp.aff_y0) / det;
    bp.ivi = -kp.aff_yx / det;
    bp.ivj = kp.aff_xx / det;
    bp.iv0 = -(-kp.aff_yx * kp.aff_x0 + kp.aff_xx * kp.aff_y0) / det;
    bp.pitch = (int)ceil(BT_W * fabs(bp.iui) + BT_R * fabs(bp.iuj)) + 3;
    bp.box_rows = (int)ceil(BT_W * fabs(bp.ivi) + BT_R * fabs(bp.ivj)) + 3;
    bp.height = bp.box_rows;
    bp.kappa = 0.0f;
    bp.kappa_d = 0.0;
    bp.wmin_di = bp.wmin_dj = 0.0;
    if (bp.iui != 0.0) {  // shear along the long sides of the region's image (slope dV/dU of the source x axis)
        const double kd = (double)(float)(bp.ivi / bp.iui);
        const double wi = bp.ivi - kd * bp.iui, wj = bp.ivj - kd * bp.iuj;  // V - kappa U per source column / row
        const int hs = (int)ceil(BT_W * fabs(wi) + BT_R * fabs(wj) + fabs(kd)) + 5;
        if (kd > 0.0 && kd <= 4.0 && hs * 20 <= bp.box_rows * 17) {  // worth the four extra instructions per pixel
            bp.kappa = (float)kd;
            bp.kappa_d = kd;
            bp.height = hs;
            bp.wmin_di = wi < 0.0 ? (double)BT_W : 0.0;
            bp.wmin_dj = wj < 0.0 ? (double)BT_R : 0.0;
        }
    }
    if (bp.pitch > 128) return (int)cudaErrorNotSupported;
    bp.emit_w = bp.pitch <= 64 ? 64 : 128;
    bp.axx = (float)kp.aff_xx;
    bp.axy = (float)kp.aff_xy;
    bp.ayx = (float)kp.aff_yx;
    bp.ayy = (float)kp.aff_yy;
    bp.umin_di = bp.iui < 0.0 ? (double)BT_W : 0.0;
    bp.umin_dj = bp.iuj < 0.0 ? (double)BT_R : 0.0;
    bp.vmin_di = bp.ivi < 0.0 ? (double)BT_W : 0.0;
    bp.vmin_dj = bp.ivj < 0.0 ? (double)BT_R : 0.0;
    bp.bu = (float)bp.iuj;
    bp.bv = (float)bp.ivj;
    const int ns = nf <= 4 ? 4 : 8;
    const size_t smem = ((((size_t)bp.pitch * bp.height + 1) * (ns * 5 + 1)) + 15) & ~(size_t)15;
    if (smem > 100 * 1024 || (int64_t)bp.pitch * bp.height >= (1 << 16)) return (int)cudaErrorNotSupported;
    // guard band of the FP32 bin coordinate: four roundings at the magnitude of the box (<= 4 * 2^-24 * extent) plus the
    // FP32 increments (extent * 2^-24 each), doubled
    const double extent = fmax(128.0, 2.0 * (double)(bp.pitch > bp.box_rows ? bp.pitch : bp.box_rows));
    bp.lim = (float)(0.5 - 12.0 * extent * 5.9604645e-8);
    // tiles that can own a pixel of the band: source bounding box of the band's centres (affine: extremes at the corners)
    double xlo = 1e300, xhi = -1e300, ylo = 1e300, yhi = -1e300;
    for (int c = 0; c < 4; ++c) {
        const double xx = (c & 1) ? (double)(kp.dst_w - 1) : 0.0, yy = (c >> 1) ? (double)(kp.row1 - 1) : (double)kp.row0;
        const double cx = kp.aff_x0 + xx * kp.aff_xx + yy * kp.aff_xy, cy = kp.aff_y0 + xx * kp.aff_yx + yy * kp.aff_yy;
        xlo = fmin(xlo, cx);
        xhi = fmax(xhi, cx);
        ylo = fmin(ylo, cy);
        yhi = fmax(yhi, cy);
    }
    const double sy0 = (double)(kp.src_y0 > 0 ? kp.src_y0 : 0);
    const double sy1 = (double)((kp.src_y0 + kp.src_rows < kp.mod_h ? kp.src_y0 + kp.src_rows : kp.mod_h) - 1);
    xlo = fmax(xlo - 1.0, 0.0);
    xhi = fmin(xhi + 1.0, (double)(kp.mod_w - 1));
    ylo = fmax(ylo - 1.0, sy0);
    yhi = fmin(yhi + 1.0, sy1);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(1, 1, kp.batch > 1 ? (unsigned)kp.batch : 1u);
    bp.tx0 = bp.ty0 = 0;
    if (xlo <= xhi && ylo <= yhi) {
        bp.tx0 = (int)floor(xlo) / BT_CW;
        bp.ty0 = (int)floor(ylo) / BT_CR;
        grid.x = (unsigned)((int)floor(xhi) / BT_CW - bp.tx0 + 1);
        grid.y = (unsigned)((int)floor(yhi) / BT_CR - bp.ty0 + 1);
        if (grid.y > 65535u) return (int)cudaErrorNotSupported;
    } else {
        bp.tx0 = bp.ty0 = 1 << 20;  // no interior pixel in this band: one CTA that emits nothing
    }
    if (dst_dtype == AAI_F32) return (int)launch_bin_to<float>(kp, bp, ns, grid, smem, st);
    return (int)launch_bin_to<uint8_t>(kp, bp, ns, grid, smem, st);
}
