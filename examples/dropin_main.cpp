// Minimal C++ host: the reference's user settings (Source.cpp:1528-1534) driven through the drop-in class.
//   g++ -std=c++17 examples/dropin_main.cpp -Larea_average_interpolation_b200 -laai_b200 -o dropin
// (CSV reading/writing of the reference's main() is row f2 of SURVEY.md §8f and not part of this example: the image
// here is synthetic.)
#include <cstdio>

#include "../area_average_interpolation_b200/csrc/aai.hpp"

int main(int argc, char **argv) {
    using namespace aai_b200;
    IMG src(911, std::vector<double>(911));
    for (size_t y = 0; y < src.size(); ++y)
        for (size_t x = 0; x < src[y].size(); ++x) src[y][x] = (double)((x * 131 + y * 71) % 4096);
    IMG dst;
    dP dstIsocenter;
    AreaAverageInterpolation aa;
    auto ret = aa.areaAverageInterpolation(src, dst, {150, 150}, {25.4, 25.4}, {455, 455}, dstIsocenter, 1.5);
    if (!ret.first) {
        std::printf("%s\nRun terminated abnormally.\n", ret.second.c_str());
        return -1;
    }
    if (argc > 1) {  // whole result as raw doubles, row-major (the parity test compares every pixel with the oracle)
        std::FILE *f = std::fopen(argv[1], "wb");
        if (!f) return -2;
        for (const auto &row : dst) std::fwrite(row.data(), sizeof(double), row.size(), f);
        std::fclose(f);
    }
    std::printf("dst %zux%zu, dstIsocenter (%g, %g), dst[79][79] = %.10g\nRun terminated correctly.\n",
                dst.empty() ? 0 : dst.front().size(), dst.size(), dstIsocenter.first, dstIsocenter.second, dst[79][79]);
    return 0;
}
