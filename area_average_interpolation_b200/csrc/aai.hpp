// aai.hpp -- C++ host mirror of the reference operator, over the C ABI (include/aai.h).
//
// Drop-in for the reference's class (Source.cpp:52-57, 584-586): same class name, same two public methods, same
// parameter list and types, same pair<bool,string> result with the reference's four messages, same ownership
// (src by value, dst cleared and resized by the callee, dstIsocenter written only after validation succeeded).
// A translation unit that used to contain the reference class can include this header instead and link
// libaai_b200.so; `main()` of Source.cpp:1434-1599 then compiles unchanged (see examples/dropin_main.cpp).
//
// Everything numeric runs on the GPU(s); without a usable device the call returns (false, "<CUDA error>") -- there is
// no CPU fallback.
#ifndef AAI_HPP_
#define AAI_HPP_

#include <string>
#include <utility>
#include <vector>

#include "../../include/aai.h"

namespace aai_b200 {

using IMG = std::vector<std::vector<double>>;       // Source.cpp:31, row-major [y][x]
using dP = std::pair<double, double>;               // Source.cpp:46, (x, y)

class AreaAverageInterpolation {
public:
    // devices: CUDA device ordinals the canvas is row-band partitioned over (empty = device 0).
    explicit AreaAverageInterpolation(std::vector<int> devices = {}, int arith = AAI_ARITH_F64)
        : devices_(std::move(devices)), arith_(arith) {}

    // Source.cpp:55-57
    std::pair<bool, std::string> areaAverageInterpolation(IMG src, IMG &dst, dP srcResolution, dP dstResolution,
                                                          dP srcIsocenter, dP &dstIsocenter, double rotationAngle) {
        return run(AAI_MODE_AREA_AVERAGE, src, dst, srcResolution, dstResolution, srcIsocenter, dstIsocenter,
                   rotationAngle);
    }
    // Source.cpp:584-586
    std::pair<bool, std::string> fastAreaAverageInterpolation(IMG src, IMG &dst, dP srcResolution, dP dstResolution,
                                                              dP srcIsocenter, dP &dstIsocenter, double rotationAngle) {
        return run(AAI_MODE_FAST, src, dst, srcResolution, dstResolution, srcIsocenter, dstIsocenter, rotationAngle);
    }

private:
    std::pair<bool, std::string> run(int mode, const IMG &src, IMG &dst, dP srcRes, dP dstRes, dP srcIso, dP &dstIso,
                                     double angle) {
        // the reference takes the width from the first row (src.front().size(), Source.cpp:150)
        const int64_t h = (int64_t)src.size();
        const int64_t w = h ? (int64_t)src.front().size() : 0;
        aai_plan plan;
        const int st = aai_plan_create(w, h, srcRes.first, srcRes.second, dstRes.first, dstRes.second, srcIso.first,
                                       srcIso.second, angle, &plan);
        if (st != AAI_OK) return {false, aai_status_string(st)};
        // beyond the reference (it reads out of bounds on ragged input): refuse it
        for (const auto &row : src)
            if ((int64_t)row.size() != w) return {false, "Ragged src array is not acceptable."};
        dstIso = {plan.dst_iso_x, plan.dst_iso_y};  // written before the main loop, like Source.cpp:181-186
        std::vector<double> flat((size_t)w * (size_t)h), out((size_t)plan.dst_w * (size_t)plan.dst_h);
        for (int64_t y = 0; y < h; ++y) std::copy(src[(size_t)y].begin(), src[(size_t)y].end(), flat.begin() + y * w);
        dst.clear();
        dst.resize((size_t)plan.dst_h);
        if (!out.empty()) {
            aai_image s{flat.data(), w * 8, w, h, 0, h, AAI_F64, 1};
            aai_image d{out.data(), plan.dst_w * 8, plan.dst_w, plan.dst_h, 0, plan.dst_h, AAI_F64, 1};
            const int rc = aai_run_host(&plan, mode, arith_, &s, &d, devices_.empty() ? nullptr : devices_.data(),
                                        (int)devices_.size());
            if (rc != AAI_OK) return {false, std::string(aai_status_string(rc)) + " " + aai_last_error()};
        }
        for (int64_t y = 0; y < plan.dst_h; ++y)
            dst[(size_t)y].assign(out.begin() + y * plan.dst_w, out.begin() + (y + 1) * plan.dst_w);
        return {true, ""};
    }

    std::vector<int> devices_;
    int arith_;
};

}  // namespace aai_b200

#endif  // AAI_HPP_
