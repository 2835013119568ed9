// Host-side parse of a Matrix Market header (plain C++: also compiled by tests/host/mtx_header_check.cxx
// on a machine without a GPU).
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace nlp {

// Header of a Matrix Market file, as readMtxHeader reads it (inc/mtx.hxx:38-55): lines that start
// with '%' are skipped, the "%%" banner names the format and the symmetry, the first other line holds
// rows, cols and the number of lines.  Returns the offset of the body (`bytes` when the text ends inside the header).
inline uint64_t mtx_header(const char* text, uint64_t bytes, bool* coordinate, bool* symmetric, uint64_t* rows, uint64_t* cols, uint64_t* size) {
  *coordinate = false; *symmetric = false; *rows = *cols = *size = 0;
  uint64_t p = 0;
  while (p < bytes) {
    uint64_t e = p;
    while (e < bytes && text[e] != '\n') ++e;
    const std::string line(text + p, text + e);
    const uint64_t next = e < bytes ? e + 1 : e;
    if (!line.empty() && line[0] == '%') {
      if (line.size() > 1 && line[1] == '%') {
        std::vector<std::string> tok;
        size_t i = 0;
        while (i < line.size()) {
          while (i < line.size() && isspace((unsigned char)line[i])) ++i;
          size_t j = i;
          while (j < line.size() && !isspace((unsigned char)line[j])) ++j;
          if (j > i) tok.push_back(line.substr(i, j - i));
          i = j;
        }
        *coordinate = tok.size() > 2 && tok[1] == "matrix" && tok[2] == "coordinate";
        *symmetric = tok.size() > 4 && (tok[4] == "symmetric" || tok[4] == "skew-symmetric");
      }
      p = next;
      continue;
    }
    unsigned long long r = 0, c = 0, n = 0;
    sscanf(line.c_str(), "%llu %llu %llu", &r, &c, &n);
    *rows = r; *cols = c; *size = n;
    return next;
  }
  return bytes;
}

}  // namespace nlp
