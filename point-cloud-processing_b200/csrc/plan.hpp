// Host-side choice of what a kNN-shaped call tries first (knn_core.cuh: SearchPlan): the level
// and the number of rings of the block, from the per-level cell counts of the index.  Shared by
// the library (query.cu) and the unit-test harness (tests/emu).  Performance only — any plan
// gives the same results.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

namespace pcpx {

struct PlanChoice
{
    int level;
    int rings;
    double expected_candidates;
};

// n points, cells[l] = occupied cells at stored level l (l = 0 .. lfine).
// For every stored level the local dimension D is read off the growth of the cell count
// (x4 per level on a surface, x8 in a volume); a block of R rings is expected to succeed when
// the ball of radius R * h holds margin * (k + 1) points, and costs the cells the k-ball touches
// times the occupancy, plus the table lookups of its rings, times a penalty for the retries a
// marginal block causes.  The cheapest pair wins.
inline PlanChoice choose_plan(uint64_t n, const uint64_t* cells, int lfine, uint32_t k,
                              double margin)
{
    PlanChoice best{0, 1, 1e300};
    double const need = margin * ((double)k + 1.0);
    for (int l = lfine; l >= 1; --l)
    {
        if (cells[l] == 0)
            continue;
        double const m = (double)n / (double)cells[l];
        double D       = cells[l - 1] > 0 ? std::log2((double)cells[l] / (double)cells[l - 1]) : 3.0;
        D              = std::min(3.0, std::max(1.0, D));
        // volume of the unit D-ball, interpolated: 2, pi, 4.19
        double const V = D <= 2.0 ? 2.0 + (D - 1.0) * (3.14159265 - 2.0)
                                  : 3.14159265 + (D - 2.0) * (4.18879 - 3.14159265);
        for (int R = 1; R <= 2; ++R)
        {
            double const inside = V * std::pow((double)R, D) * m; // points within R * h
            // A block whose ball is expected to hold somewhat fewer than `need` points is still
            // tried — a failed attempt falls back to the retry kernel — at a price that grows with
            // the shortfall (measured, 10 M-point plane: k = 30 takes 7.5 ms with one ring and
            // its retries against 9.7 ms with two rings, k = 28 6.5 ms against 9.0 ms).
            double const fill = inside / need;
            if (fill < 0.7)
                continue;
            double const retry_penalty = 1.0 + 3.0 * std::max(0.0, 1.0 - fill);
            double const rho  = std::min((double)R, std::pow(((double)k + 1.0) / (V * m), 1.0 / D));
            // the lookups of the rings: 27 for one ring; the second ring walks 98 more offsets
            // whether or not they survive the bound test
            double const cost =
                (m * std::pow(2.0 * rho + 1.0, D) + (R == 1 ? 13.5 : 150.0)) * retry_penalty;
            if (cost < best.expected_candidates)
                best = PlanChoice{l, R, cost};
        }
    }
    return best; // level 0 (the whole cloud in one cell) when nothing finer is adequate
}

} // namespace pcpx
