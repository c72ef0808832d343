// MSM launch plan (host + device POD) and the window-size cost model.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>

namespace b200 {

constexpr uint32_t kOvfTaskPoints = 128;   // points per overflow task (their latency is serial)

struct MsmPlan {
  uint64_t n;          // number of (point, scalar) pairs
  int c;               // window width in bits
  int nwin;            // ceil((scalar_bits + 1) / c)
  uint32_t nb;         // buckets per window = 2^(c-1)
  uint32_t task;       // max points a single thread accumulates for one bucket
  uint32_t group;      // buckets per running-sum group
  uint32_t max_ovf;    // capacity of the overflow task list
};

inline MsmPlan make_msm_plan(uint64_t n, int scalar_bits, int c_override) {
  MsmPlan pl{};
  pl.n = n;
  int best_c = 2;
  double best = 1e300;
  int cmax = 16;
  for (int c = 2; c <= cmax; c++) {
    int nwin = (scalar_bits + 1 + c - 1) / c;
    double nb = std::ldexp(1.0, c - 1);
    // mixed adds for accumulation + ~3x weight for the (full-add, low-parallelism) bucket reduction
    double cost = (double)n * nwin * 1.0 + nwin * nb * 2.0 * 3.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  pl.c = c_override > 0 ? c_override : best_c;
  pl.nwin = (scalar_bits + 1 + pl.c - 1) / pl.c;
  pl.nb = 1u << (pl.c - 1);
  uint64_t avg = n / pl.nb + 1;
  pl.task = (uint32_t)std::max<uint64_t>(256, 8 * avg);
  uint64_t total_b = (uint64_t)pl.nwin * pl.nb;
  uint32_t g = 1;
  while (g * 2 <= 64 && (uint64_t)g * 2 * 16384 <= total_b) g *= 2;
  if (g < 4) g = std::min<uint32_t>(4, pl.nb);
  pl.group = std::min<uint32_t>(g, pl.nb);
  pl.max_ovf = (uint32_t)((n * (uint64_t)pl.nwin) / kOvfTaskPoints + 1);
  return pl;
}

}  // namespace b200
