// MSM launch plan (host + device POD) and the window-size cost model.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b200 {

constexpr uint32_t kOvfTaskPoints = 32;    // points per overflow task (a chain of mixed adds is pure latency, ~10 us each)
constexpr int kMaxSets = 4;                // bucket sets one sorting pass can feed (table mode, shared scalar vector)

// Two modes:
//  * windowed (bwin == nwin): classic Pippenger, one bucket set per scalar window, Horner at the end.
//  * table (bwin == 1): the base set carries precomputed multiples T_j[i] = 2^(c j) P_i resident in
//    HBM, so every window's digit adds into ONE shared bucket set - no Horner, nwin times fewer
//    buckets to reduce, which lets c grow (fewer point additions).  Several base sets that are multiplied
//    by the SAME scalar vector (the A / B / K keys of a Groth16 proof all take the wire vector) share one
//    digit/sort pass: bwin == nsets bucket sets, one per base set, each with its own index map.
//    HBM budget knob: with table stride s only every s-th table is kept (T_q = 2^(c s q) P, ceil(nwin / s)
//    tables); digit window w = q s + r reads table q and adds into bucket set r of its base set (s bucket
//    sets per base set, bwin == nsets * s), and the set's result is sum_r 2^(c r) S_r - a Horner of (s - 1) c
//    doublings of ONE point.  s = 1 is the full table mode, s = nwin degenerates to windowed Pippenger.
struct MsmPlan {
  uint64_t n;          // number of scalars
  int c;               // window width in bits
  int nwin;            // digit windows = ceil((scalar_bits + 1) / c)
  int bwin;            // bucket arrays: nwin (windowed) or the number of base sets (table mode)
  int table;           // 1 = table mode
  uint32_t batch_n;    // > 0: batched mode - `bwin` scalar vectors of batch_n scalars each share ONE base set (and its
                       // index map); scalar i belongs to vector i / batch_n and multiplies base (i % batch_n)
  int pre;             // affine pre-reduction levels (msm_pre.cuh): bucket segments are padded to multiples of 2^pre
  uint32_t nb;         // buckets per bucket window = 2^(c-1)
  uint32_t task;       // max points a single thread accumulates for one bucket
  uint32_t task_min;   // lower bound of the load-adaptive cap
  uint32_t ovf_task;   // points per overflow task (the remainder of an oversized bucket)
  uint32_t group;      // buckets per running-sum group
  uint32_t max_ovf;    // capacity of the overflow task list
  uint64_t npts;       // table mode: points per table (index of digit window w is (w / tstride)*npts + i)
  uint64_t stride;     // sorted-index entries reserved per bucket window
  int tstride;         // table mode: table stride s (>= 1); bucket set of (base set j, window w) is j*s + w % s
};

inline int msm_nwin(int scalar_bits, int c) { return (scalar_bits + 1 + c - 1) / c; }
inline int msm_ntables(int nwin, int tstride) { return (nwin + tstride - 1) / tstride; }

// window width for a table-mode base set of npts points (fixed when the tables are built); with table stride s the
// bucket reduction is s times as large
inline int msm_table_window(uint64_t npts, int scalar_bits, int tstride = 1) {
  // small base sets (e.g. the 4096-point KZG SRS) are latency-bound: every sequential point addition
  // costs ~8 us on an almost empty GPU, so take the smallest window that leaves <= ~4 points per bucket
  if (npts <= (1u << 16)) {
    for (int c = 8; c <= 16; c++) {
      double nb = std::ldexp(1.0, c - 1);
      if ((double)npts * msm_nwin(scalar_bits, c) / nb <= 4.0) return c;
    }
    return 16;
  }
  int best_c = 4;
  double best = 1e300;
  for (int c = 4; c <= 22; c++) {
    int nwin = msm_nwin(scalar_bits, c);
    double nb = std::ldexp(1.0, c - 1);
    const int s = tstride < nwin ? tstride : nwin;
    double cost = (double)npts * nwin + nb * 2.0 * 4.0 * s + 3000.0 * nwin;   // adds + bucket reduction + fixed
    if ((double)msm_ntables(nwin, s) * (double)npts >= 2147483648.0) continue;   // index must fit 31 bits
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

inline void msm_plan_finish(MsmPlan& pl) {
  pl.nb = 1u << (pl.c - 1);
  const uint64_t adds = pl.n * (uint64_t)pl.nwin;
  const uint64_t total_b = (uint64_t)pl.bwin * pl.nb;
  uint64_t avg = adds / (pl.table ? (uint64_t)pl.nb * pl.tstride : total_b) + 1;
  // large MSMs: size-sorted scheduling hides long buckets, keep overflow rare.  small MSMs: the longest
  // per-thread chain IS the latency, so cap it hard and split the rest into short parallel tasks.
  const bool small = adds < (1u << 21);
  // `task` is the upper cap; the kernel lowers it to 4x the measured mean load (never below task_min)
  pl.task = small ? (uint32_t)std::max<uint64_t>(16, 4 * avg) : (uint32_t)std::max<uint64_t>(256, 8 * avg);
  pl.task_min = small ? 16 : 64;
  pl.ovf_task = small ? (uint32_t)std::min<uint64_t>(kOvfTaskPoints, std::max<uint64_t>(16, 2 * avg)) : kOvfTaskPoints;
  // running-sum groups: short groups keep the (latency-bound) bucket reduction shallow
  uint32_t g = 1;
  static const uint32_t gmax = [] {
    const char* e = std::getenv("B200_MSM_GROUP_MAX");   // tuning knob (default 32)
    return e ? (uint32_t)std::atoi(e) : 32u;
  }();
  while (g * 2 <= gmax && (uint64_t)g * 2 * 16384 <= total_b) g *= 2;
  if (g < 4) g = std::min<uint32_t>(4, pl.nb);
  pl.group = std::min<uint32_t>(g, pl.nb);
  // every overflowing bucket holds > task_min points and every overflow task but the last is full
  pl.max_ovf = (uint32_t)((adds / pl.ovf_task + adds / pl.task_min + 2) * (pl.table ? pl.bwin / pl.tstride : 1));
  // table mode: a bucket set receives at most ceil(nwin / s) digits per scalar
  pl.stride = pl.table ? pl.n * (uint64_t)msm_ntables(pl.nwin, pl.tstride) + ((uint64_t)pl.nb << pl.pre) : pl.n;
}

inline MsmPlan make_msm_plan(uint64_t n, int scalar_bits, int c_override) {
  MsmPlan pl{};
  pl.n = n;
  int best_c = 2;
  double best = 1e300;
  for (int c = 2; c <= 16; c++) {
    int nwin = msm_nwin(scalar_bits, c);
    double nb = std::ldexp(1.0, c - 1);
    // mixed adds for accumulation + ~3x weight for the (full-add, low-parallelism) bucket reduction
    double cost = (double)n * nwin * 1.0 + nwin * nb * 2.0 * 3.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  pl.c = c_override > 0 ? c_override : best_c;
  pl.nwin = msm_nwin(scalar_bits, pl.c);
  pl.bwin = pl.nwin;
  pl.table = 0;
  pl.npts = 0;
  pl.tstride = 1;
  msm_plan_finish(pl);
  return pl;
}

inline MsmPlan make_msm_plan_table(uint64_t n, int scalar_bits, int c, uint64_t npts, int nsets = 1, int pre = 0,
                                   int tstride = 1) {
  MsmPlan pl{};
  pl.n = n;
  pl.pre = (nsets == 1 && tstride == 1) ? pre : 0;
  pl.c = c;
  pl.nwin = msm_nwin(scalar_bits, c);
  pl.tstride = tstride < 1 ? 1 : (tstride > pl.nwin ? pl.nwin : tstride);
  pl.bwin = nsets * pl.tstride;
  pl.table = 1;
  pl.npts = npts;
  msm_plan_finish(pl);
  return pl;
}

}  // namespace b200
