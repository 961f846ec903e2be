// aai_csv.hpp -- CSV image reader / writer with the reference's file conventions.
//
// Written from the observable behaviour of the reference's `csvRead` / `csvWrite` / `splitPath` lambdas
// (Source.cpp:1437-1515), not from their text:
//   * a file is a sequence of lines; a line is a sequence of comma-separated fields; a field is read as the longest
//     numeric prefix a C++ `stod` would accept (leading white space allowed); a field with no numeric prefix is
//     skipped, not an error (Source.cpp:1453-1468);
//   * every line becomes one image row;
//   * the writer prints each value with the stream's default format (6 significant digits, %g style) separated by
//     commas, one line per row (Source.cpp:1505-1512) -- the output is therefore lossy by design;
//   * the output file name is `<dir><base>_mod<ext>` (Source.cpp:1591).
// Differences, all on inputs where the reference has undefined behaviour: rows of different length and blank lines
// are reported as errors (the reference reads out of bounds, Source.cpp:1486-1488).
#ifndef AAI_CSV_HPP_
#define AAI_CSV_HPP_

#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace aai_b200 {

using IMG = std::vector<std::vector<double>>;

struct PathParts {
    std::string dir, base, ext;  // "a/b/", "name", ".csv"
};

inline PathParts split_path(const std::string &full) {
    PathParts p;
    const size_t dot = full.rfind('.');
    size_t sep = full.rfind('\\');
    if (sep == std::string::npos) sep = full.rfind('/');
    const size_t start = sep == std::string::npos ? 0 : sep + 1;
    p.ext = dot == std::string::npos ? "" : full.substr(dot);
    p.base = full.substr(start, dot == std::string::npos ? std::string::npos : dot - start);
    p.dir = full.substr(0, start);
    return p;
}

// Returns "" on success, else an error message.
inline std::string csv_read(const std::string &path, IMG &image) {
    std::ifstream in(path);
    if (!in) return "Failed to read csv file.";
    image.clear();
    std::string line;
    size_t width = 0;
    while (std::getline(in, line)) {
        std::vector<double> row;
        size_t pos = 0;
        while (pos <= line.size()) {
            size_t comma = line.find(',', pos);
            if (comma == std::string::npos) comma = line.size();
            const std::string field = line.substr(pos, comma - pos);
            char *end = nullptr;
            errno = 0;
            const double v = std::strtod(field.c_str(), &end);
            if (end != field.c_str()) row.push_back(v);  // no numeric prefix: the field is ignored
            pos = comma + 1;
        }
        if (row.empty()) return "Blank line in csv file (the reference's behaviour is undefined here).";
        if (image.empty()) width = row.size();
        if (row.size() != width) return "Rows of different length in csv file (the reference reads out of bounds here).";
        image.push_back(std::move(row));
    }
    return "";
}

inline std::string csv_write(const std::string &path, const IMG &image) {
    std::ofstream out(path);
    if (!out) return "Failed to write csv file.";
    if (image.empty()) return "There is no data in src array.";
    const size_t width = image.front().size();
    for (const auto &row : image) {
        for (size_t j = 0; j < width; ++j) {
            out << row[j];  // default stream format: 6 significant digits
            if (j + 1 < width) out << ",";
        }
        out << std::endl;
    }
    return "";
}

}  // namespace aai_b200

#endif  // AAI_CSV_HPP_
