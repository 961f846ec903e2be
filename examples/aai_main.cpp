// aai_main -- the reference's command-line driver (Source.cpp:1434-1599) on top of the B200 library.
//
//   aai_main [input.csv [srcRes dstRes isoX isoY angleDeg mode]]
//
// Without arguments it runs the reference's hard-coded user settings (Source.cpp:1528-1534): Test_film_dose.csv,
// 150 -> 25.4 dpi, isocentre (455,455), 1.5 degrees, mode 2 (fast area average); it prints the same progress lines and
// writes `<base>_mod.csv` next to the input, in the reference's 6-significant-digit format.
//   g++ -std=c++17 examples/aai_main.cpp -Larea_average_interpolation_b200 -laai_b200 -o aai_main
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "../area_average_interpolation_b200/csrc/aai.hpp"
#include "../area_average_interpolation_b200/csrc/aai_csv.hpp"

int main(int argc, char **argv) {
    using namespace aai_b200;
    std::string inputPath = "Test_film_dose.csv";
    double srcRes = 150, dstRes = 25.4, isoX = 455, isoY = 455, rotationAngle = 1.5;
    int interpolationMode = 2;
    if (argc > 1) inputPath = argv[1];
    if (argc > 7) {
        srcRes = std::atof(argv[2]);
        dstRes = std::atof(argv[3]);
        isoX = std::atof(argv[4]);
        isoY = std::atof(argv[5]);
        rotationAngle = std::atof(argv[6]);
        interpolationMode = std::atoi(argv[7]);
    }
    const PathParts parts = split_path(inputPath);
    if (parts.ext != ".csv" && parts.ext != ".CSV") {
        std::cout << "As for the image format, only csv format can be used." << std::endl;
        std::cout << "Run terminated abnormally." << std::endl;
        return -1;
    }
    IMG src, dst;
    std::string err = csv_read(inputPath, src);
    if (!err.empty()) {
        std::cout << err << std::endl << "Run terminated abnormally." << std::endl;
        return -1;
    }
    AreaAverageInterpolation aa;
    dP dstIsocenter;
    std::pair<bool, std::string> ret;
    const auto start = std::chrono::system_clock::now();
    switch (interpolationMode) {
        case 1:
            ret = aa.areaAverageInterpolation(src, dst, {srcRes, srcRes}, {dstRes, dstRes}, {isoX, isoY}, dstIsocenter,
                                              rotationAngle);
            break;
        case 2:
            ret = aa.fastAreaAverageInterpolation(src, dst, {srcRes, srcRes}, {dstRes, dstRes}, {isoX, isoY},
                                                  dstIsocenter, rotationAngle);
            break;
        default:
            std::cout << "Invalid interpolation mode is selected." << std::endl;
            std::cout << "Interpolation mode should be 1 or 2." << std::endl;
            std::cout << "Run terminated abnormally." << std::endl;
            return -1;
    }
    const auto end = std::chrono::system_clock::now();
    std::cout << "Calculation time : "
              << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() / 1000.0 << " [ms]" << std::endl;
    if (!ret.first) {
        std::cout << ret.second << std::endl << "Run terminated abnormally." << std::endl;
        return -1;
    }
    err = csv_write(parts.dir + parts.base + "_mod" + parts.ext, dst);
    if (!err.empty()) {
        std::cout << err << std::endl << "Run terminated abnormally." << std::endl;
        return -1;
    }
    std::cout << "Run terminated correctly." << std::endl;
    return 0;
}
