// TEST-ONLY harness around csrc/aai_csv.hpp:  csv_host read <in.csv> <out.f64>   |   csv_host write <in.f64> <w> <h> <out.csv>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "../area_average_interpolation_b200/csrc/aai_csv.hpp"

int main(int argc, char **argv) {
    using namespace aai_b200;
    if (argc >= 4 && !std::strcmp(argv[1], "read")) {
        IMG img;
        const std::string err = csv_read(argv[2], img);
        if (!err.empty()) {
            std::cout << err << std::endl;
            return 2;
        }
        FILE *f = std::fopen(argv[3], "wb");
        for (auto &row : img) std::fwrite(row.data(), sizeof(double), row.size(), f);
        std::fclose(f);
        std::cout << img.front().size() << " " << img.size() << std::endl;
        return 0;
    }
    if (argc >= 6 && !std::strcmp(argv[1], "write")) {
        const size_t w = std::strtoul(argv[3], nullptr, 10), h = std::strtoul(argv[4], nullptr, 10);
        IMG img(h, std::vector<double>(w));
        FILE *f = std::fopen(argv[2], "rb");
        for (auto &row : img)
            if (std::fread(row.data(), sizeof(double), w, f) != w) return 3;
        std::fclose(f);
        const std::string err = csv_write(argv[5], img);
        if (!err.empty()) {
            std::cout << err << std::endl;
            return 2;
        }
        return 0;
    }
    if (argc >= 3 && !std::strcmp(argv[1], "split")) {
        const PathParts p = split_path(argv[2]);
        std::cout << p.dir << "|" << p.base << "|" << p.ext << std::endl;
        return 0;
    }
    return 1;
}
