// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product path.
//
// Flat C bridge over the UNMODIFIED reference headers, compiled where they lie under
// /root/reference/include (no reference source is copied into this repository) by
// oracle/Makefile into oracle/_ref/libpcp_ref.so. It is the strongest pin the parity tests
// have: the reference's own
//   pcp::basic_linked_octree_t::nearest_neighbours   include/pcp/octree/linked_octree.hpp:245-254
//   pcp::basic_linked_octree_t::range_search         include/pcp/octree/linked_octree.hpp:264-276
//   pcp::basic_linked_kdtree_t::nearest_neighbours   include/pcp/kdtree/linked_kdtree.hpp:200-263
//   pcp::basic_linked_kdtree_t::range_search         include/pcp/kdtree/linked_kdtree.hpp:270-277
//   pcp::algorithm::average_distances_to_neighbors   include/pcp/algorithm/average_distance_to_neighbors.hpp:39-73
// run here on flat float buffers. pcp::estimate_normal cannot be compiled (needs Eigen 3.3.8,
// fetched by the reference's CMake and absent from /root/reference and from this image);
// the restatement in oracle/pcp_oracle.c covers it.
//
// Build flags that matter (see oracle/Makefile): -ffp-contract=off and NO -march=native so
// that squared_distance is evaluated without FMA, exactly as a baseline x86-64 build of the
// reference would.
//
// `std::execution::par` is serial in this image (libstdc++'s PSTL backend needs TBB), so the
// batched entry points partition the queries statically over `nthreads` std::threads, which
// is what the policy would do with a working backend; every query method used is `const`.
#include <algorithm> // must precede the octree headers (linked_octree_iterator.hpp uses std::find_if)
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <pcp/algorithm/average_distance_to_neighbors.hpp>
#include <pcp/common/points/point.hpp>
#include <pcp/common/points/point_view.hpp>
#include <pcp/common/sphere.hpp>
#include <pcp/kdtree/kdtree.hpp>
#include <pcp/octree/octree.hpp>

namespace {

struct view_map_t
{
    pcp::point_view_t operator()(pcp::point_view_t const& p) const { return p; }
};

struct coord_map_t
{
    std::array<float, 3> operator()(pcp::point_view_t const& p) const
    {
        return std::array<float, 3>{p.x(), p.y(), p.z()};
    }
};

using octree_t = pcp::basic_linked_octree_t<pcp::point_view_t>;
using kdtree_t = pcp::basic_linked_kdtree_t<pcp::point_view_t, 3u, coord_map_t>;

struct ref_cloud_t
{
    std::vector<pcp::point_t> points;
    std::vector<pcp::point_view_t> views;
    std::unique_ptr<octree_t> octree;
    std::unique_ptr<kdtree_t> kdtree;

    std::int64_t index_of(pcp::point_view_t const& v) const
    {
        return static_cast<std::int64_t>(v.point() - points.data());
    }
};

template <class F>
void parallel_for(std::size_t n, int nthreads, F&& f)
{
    if (nthreads <= 1 || n < 2)
    {
        for (std::size_t i = 0; i < n; ++i)
            f(i);
        return;
    }
    std::vector<std::thread> pool;
    std::size_t const T = static_cast<std::size_t>(nthreads);
    for (std::size_t t = 0; t < T; ++t)
    {
        std::size_t const b = n * t / T, e = n * (t + 1) / T;
        pool.emplace_back([b, e, &f]() {
            for (std::size_t i = b; i < e; ++i)
                f(i);
        });
    }
    for (auto& th : pool)
        th.join();
}

} // namespace

extern "C" {

// which = 0 → octree (auto bbox unless bbox6 != NULL), 1 → kd-tree, 2 → both
void* ref_cloud_create(
    float const* xyz,
    std::size_t n,
    int which,
    float const* bbox6,
    std::uint32_t node_capacity,
    std::uint32_t max_depth,
    int kd_adaptive_depth)
{
    auto* c = new ref_cloud_t{};
    c->points.reserve(n);
    for (std::size_t i = 0; i < n; ++i)
        c->points.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    c->views.reserve(n);
    for (std::size_t i = 0; i < n; ++i)
        c->views.emplace_back(&c->points[i]);

    if (which == 0 || which == 2)
    {
        if (bbox6 != nullptr)
        {
            pcp::octree_parameters_t<pcp::point_t> params;
            if (node_capacity)
                params.node_capacity = node_capacity;
            if (max_depth)
                params.max_depth = static_cast<std::uint8_t>(max_depth);
            params.voxel_grid = pcp::axis_aligned_bounding_box_t<pcp::point_t>{
                pcp::point_t{bbox6[0], bbox6[1], bbox6[2]},
                pcp::point_t{bbox6[3], bbox6[4], bbox6[5]}};
            c->octree = std::make_unique<octree_t>(
                c->views.begin(),
                c->views.end(),
                view_map_t{},
                params);
        }
        else
        {
            c->octree = std::make_unique<octree_t>(c->views.begin(), c->views.end(), view_map_t{});
        }
    }
    if (which == 1 || which == 2)
    {
        pcp::kdtree::construction_params_t params;
        params.compute_max_depth = kd_adaptive_depth != 0;
        if (!kd_adaptive_depth && max_depth)
            params.max_depth = max_depth;
        c->kdtree =
            std::make_unique<kdtree_t>(c->views.begin(), c->views.end(), coord_map_t{}, params);
    }
    return c;
}

void ref_cloud_destroy(void* h) { delete static_cast<ref_cloud_t*>(h); }

std::size_t ref_octree_size(void* h)
{
    auto* c = static_cast<ref_cloud_t*>(h);
    return c->octree ? c->octree->size() : 0u;
}

void ref_octree_bbox(void* h, float* out6)
{
    auto* c       = static_cast<ref_cloud_t*>(h);
    auto const& b = c->octree->voxel_grid();
    out6[0] = b.min.x(), out6[1] = b.min.y(), out6[2] = b.min.z();
    out6[3] = b.max.x(), out6[4] = b.max.y(), out6[5] = b.max.z();
}

// Raw reference output, in the reference's own order. out_idx is nq×k, padded with -1.
// tree = 0 octree, 1 kd-tree. queries == NULL → the cloud's own points are the targets.
void ref_knn(
    void* h,
    int tree,
    float const* queries,
    std::size_t nq,
    std::size_t k,
    double eps,
    std::int64_t* out_idx,
    std::uint32_t* out_count,
    int nthreads)
{
    auto* c = static_cast<ref_cloud_t*>(h);
    parallel_for(nq, nthreads, [&](std::size_t i) {
        pcp::point_t const t =
            queries ? pcp::point_t{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]} :
                      c->points[i];
        std::vector<pcp::point_view_t> nn;
        if (tree == 0)
            nn = c->octree->nearest_neighbours(t, k, view_map_t{}, eps);
        else
            nn = c->kdtree->nearest_neighbours(
                std::array<float, 3>{t.x(), t.y(), t.z()},
                k,
                static_cast<float>(eps));
        for (std::size_t j = 0; j < k; ++j)
            out_idx[i * k + j] = j < nn.size() ? c->index_of(nn[j]) : -1;
        if (out_count)
            out_count[i] = static_cast<std::uint32_t>(nn.size());
    });
}

// Tie-aware reference kNN: the order among bit-equal fp32 distances in the
// reference is an artefact of std::priority_queue, so ask the reference for k+extra, recompute
// the reference's own squared_distance, order by (d², original index), and enlarge the request
// while the boundary distance is still tied. Output: nq×k indices (pad -1), nq×k d² (pad +inf).
void ref_knn_tie_aware(
    void* h,
    int tree,
    float const* queries,
    std::size_t nq,
    std::size_t k,
    double eps,
    std::int64_t* out_idx,
    float* out_d2,
    std::uint32_t* out_count,
    int nthreads)
{
    auto* c              = static_cast<ref_cloud_t*>(h);
    std::size_t const n  = c->points.size();
    parallel_for(nq, nthreads, [&](std::size_t i) {
        pcp::point_t const t =
            queries ? pcp::point_t{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]} :
                      c->points[i];
        std::size_t ask = k + 8;
        std::vector<std::pair<float, std::int64_t>> cand;
        for (;;)
        {
            std::vector<pcp::point_view_t> nn;
            if (tree == 0)
                nn = c->octree->nearest_neighbours(t, ask, view_map_t{}, eps);
            else
                nn = c->kdtree->nearest_neighbours(
                    std::array<float, 3>{t.x(), t.y(), t.z()},
                    ask,
                    static_cast<float>(eps));
            cand.clear();
            for (auto const& v : nn)
                cand.emplace_back(pcp::common::squared_distance(t, v), c->index_of(v));
            std::sort(cand.begin(), cand.end());
            bool const exhausted = nn.size() < ask || ask >= n;
            // safe when the furthest returned distance is strictly beyond the k-th kept one
            if (exhausted || cand.size() <= k || cand.back().first > cand[k - 1].first)
                break;
            ask *= 2;
        }
        std::size_t const m = std::min(k, cand.size());
        for (std::size_t j = 0; j < k; ++j)
        {
            out_idx[i * k + j] = j < m ? cand[j].second : -1;
            if (out_d2)
                out_d2[i * k + j] = j < m ? cand[j].first : INFINITY;
        }
        if (out_count)
            out_count[i] = static_cast<std::uint32_t>(m);
    });
}

// Sphere range search. radii == NULL → every query uses r. Counts always; when
// out_offsets/out_idx are given (two-call protocol: first call with out_idx == NULL fills
// counts, caller prefix-sums into offsets) indices are written sorted ascending per query.
void ref_radius(
    void* h,
    int tree,
    float const* queries,
    std::size_t nq,
    float const* radii,
    float r,
    std::uint32_t* out_count,
    std::uint64_t const* offsets,
    std::int64_t* out_idx,
    int nthreads)
{
    auto* c = static_cast<ref_cloud_t*>(h);
    parallel_for(nq, nthreads, [&](std::size_t i) {
        pcp::point_t const t =
            queries ? pcp::point_t{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]} :
                      c->points[i];
        float const ri = radii ? radii[i] : r;
        std::vector<pcp::point_view_t> in;
        if (tree == 0)
        {
            pcp::sphere_t<pcp::point_t> s;
            s.position = t;
            s.radius   = ri;
            in         = c->octree->range_search(s, view_map_t{});
        }
        else
        {
            pcp::sphere_a<float> s{{t.x(), t.y(), t.z()}, ri};
            in = c->kdtree->range_search(s);
        }
        if (out_count)
            out_count[i] = static_cast<std::uint32_t>(in.size());
        if (out_idx && offsets)
        {
            std::vector<std::int64_t> ids;
            ids.reserve(in.size());
            for (auto const& v : in)
                ids.push_back(c->index_of(v));
            std::sort(ids.begin(), ids.end());
            std::copy(ids.begin(), ids.end(), out_idx + offsets[i]);
        }
    });
}

// pcp::algorithm::average_distances_to_neighbors over the cloud's own points with the octree
// (tree = 0) or kd-tree (tree = 1) as KnnMap. out_means has n entries; returns the reference's
// own average_distance_to_neighbors result (fp32 std::reduce, sequential here).
float ref_average_distance_to_neighbors(void* h, int tree, std::size_t k, float* out_means)
{
    auto* c              = static_cast<ref_cloud_t*>(h);
    auto const point_map = [](pcp::point_view_t const& v) {
        return pcp::point_t{v};
    };
    auto const knn_map = [&](pcp::point_view_t const& v) {
        if (tree == 0)
            return c->octree->nearest_neighbours(v, k, view_map_t{});
        return c->kdtree->nearest_neighbours(v, k);
    };
    std::vector<float> means = pcp::algorithm::average_distances_to_neighbors(
        c->views.begin(),
        c->views.end(),
        point_map,
        knn_map);
    if (out_means)
        std::copy(means.begin(), means.end(), out_means);
    return pcp::algorithm::average_distance_to_neighbors(
        c->views.begin(),
        c->views.end(),
        point_map,
        knn_map);
}

// The per-element body of pcp::algorithm::estimate_normals (algorithm/estimate_normals.hpp:80-90)
// for a sample of the cloud's own points: knn(v) through the reference's tree, then
// pcp::estimate_normal.  The latter needs Eigen (absent), so the restatement in
// oracle/pcp_oracle.c (oracle_estimate_normal, common/normals/normal_estimation.hpp:41-77) is
// linked in for that one call; the neighbour search is the reference's own code.
void oracle_estimate_normal(const float* pts, std::size_t n, float* out3, float* out_gap);

void ref_estimate_normals_sample(
    void* h,
    int tree,
    std::uint32_t const* sample,
    std::size_t ns,
    std::size_t k,
    float* out_normals,
    int nthreads)
{
    auto* c = static_cast<ref_cloud_t*>(h);
    parallel_for(ns, nthreads, [&](std::size_t i) {
        pcp::point_view_t const& v = c->views[sample[i]];
        std::vector<pcp::point_view_t> nn;
        if (tree == 0)
            nn = c->octree->nearest_neighbours(v, k, view_map_t{});
        else
            nn = c->kdtree->nearest_neighbours(v, k);
        std::vector<float> pts(3 * nn.size());
        for (std::size_t j = 0; j < nn.size(); ++j)
            pts[3 * j] = nn[j].x(), pts[3 * j + 1] = nn[j].y(), pts[3 * j + 2] = nn[j].z();
        oracle_estimate_normal(pts.data(), nn.size(), out_normals + 3 * i, nullptr);
    });
}

} // extern "C"
