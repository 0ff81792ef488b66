// csv_tool.cpp — writes one trajectory in the reference's result format (src/ilqr_core.cpp:414-431) with the host
// layer's own writer, from raw arrays:   csv_tool T n m in.bin out.csv    (in.bin: xs [T+1][n] then us [T][m], f64).
// Needs no GPU; the CPU tests use it to compare the writer byte for byte with the reference's file.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ilqr.h"

int main(int argc, char **argv) {
  if (argc != 6) return 2;
  const int T = atoi(argv[1]), n = atoi(argv[2]), m = atoi(argv[3]);
  std::vector<double> xs((size_t)(T + 1) * n), us((size_t)T * m);
  FILE *in = fopen(argv[4], "rb");
  if (!in) return 3;
  const bool ok = fread(xs.data(), sizeof(double), xs.size(), in) == xs.size() && fread(us.data(), sizeof(double), us.size(), in) == us.size();
  fclose(in);
  if (!ok) return 4;
  FILE *out = fopen(argv[5], "w");
  if (!out) return 5;
  ilqr_write_csv(out, T, n, m, xs.data(), us.data());
  fclose(out);
  return 0;
}
