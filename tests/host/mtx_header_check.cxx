// CPU check of the Matrix Market header parse used by nlp_ingest_mtx (csrc/mtx_header.hpp) against
// what the reference's readMtxHeader (inc/mtx.hxx:38-55) accepts.
#include <cstdio>
#include <cstring>
#include <string>
#include "mtx_header.hpp"

static int fails = 0;

static void expect(const char* name, const std::string& text, bool coordinate, bool symmetric, uint64_t rows, uint64_t cols,
                   uint64_t size, const char* body) {
  bool c = false, s = false;
  uint64_t r = 0, k = 0, n = 0;
  const uint64_t at = nlp::mtx_header(text.data(), text.size(), &c, &s, &r, &k, &n);
  const std::string rest = text.substr((size_t)at);
  if (c != coordinate || s != symmetric || r != rows || k != cols || n != size || rest != body) {
    printf("FAIL %s: coordinate %d symmetric %d rows %llu cols %llu size %llu body [%s]\n", name, (int)c, (int)s,
           (unsigned long long)r, (unsigned long long)k, (unsigned long long)n, rest.c_str());
    ++fails;
  }
}

int main() {
  expect("general", "%%MatrixMarket matrix coordinate real general\n% c\n3 4 2\n1 2 1\n2 3 1\n", true, false, 3, 4, 2, "1 2 1\n2 3 1\n");
  expect("symmetric", "%%MatrixMarket matrix coordinate pattern symmetric\n5 5 1\n2 1\n", true, true, 5, 5, 1, "2 1\n");
  expect("skew", "%%MatrixMarket matrix coordinate real skew-symmetric\n5 5 0\n", true, true, 5, 5, 0, "");
  expect("crlf", "%%MatrixMarket matrix coordinate integer symmetric\r\n%x\r\n7 6 1\r\n2 1 5\r\n", true, true, 7, 6, 1, "2 1 5\r\n");
  expect("array", "%%MatrixMarket matrix array real general\n3 3\n1\n", false, false, 3, 3, 0, "1\n");
  expect("no banner", "% only a comment\n2 2 1\n1 2\n", false, false, 2, 2, 1, "1 2\n");
  expect("no newline after the size line", "%%MatrixMarket matrix coordinate real general\n2 2 0", true, false, 2, 2, 0, "");
  expect("comments only", "%%MatrixMarket matrix coordinate real general\n% a\n% b\n", true, false, 0, 0, 0, "");
  expect("extra blanks", "%%MatrixMarket   matrix\tcoordinate  real   general  \n  10   12   3  \n1 1 1\n", true, false, 10, 12, 3, "1 1 1\n");
  if (fails) return 1;
  printf("mtx_header_check ok\n");
  return 0;
}
