// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product path.
//
// Flat C bridge over the UNMODIFIED reference for normal orientation,
// compiled where the headers lie under /root/reference/include by oracle/Makefile into
// oracle/_ref/libpcp_ref_orient.so:
//   pcp::algorithm::propagate_normal_orientations   include/pcp/algorithm/estimate_normals.hpp:187-302
//   pcp::graph::directed_knn_graph                  include/pcp/graph/knn_adjacency_list.hpp:117-156
//   pcp::graph::breadth_first_search                include/pcp/graph/search.hpp:41-85
//   pcp::basic_linked_kdtree_t::nearest_neighbours  include/pcp/kdtree/linked_kdtree.hpp:200-263
// called the way examples/normals_estimation.cpp:69-117 calls them.
//
// estimate_normals.hpp includes pcp/common/normals/normal_estimation.hpp (concrete Eigen types,
// Eigen 3.3.8 absent here); oracle/ref_shim_orient/ shadows exactly that one header with a
// declaration of estimate_normal, which propagate_normal_orientations never calls.
//
// kNN lists may be passed in (knn != NULL: n x k indices, -1 = none) so that ties in the
// reference's own neighbour order cannot blur the comparison; knn == NULL uses the
// reference's kd-tree.
#include <algorithm>
#include <array>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <vector>

#include <pcp/algorithm/estimate_normals.hpp>
#include <pcp/common/normals/normal.hpp>
#include <pcp/common/points/point.hpp>
#include <pcp/kdtree/kdtree.hpp>

extern "C" void ref_propagate_normal_orientations(const float* xyz, std::size_t n,
                                                  const std::int64_t* knn, std::size_t k,
                                                  float* normals)
{
    std::vector<pcp::point_t> points(n);
    std::vector<pcp::normal_t> nrm(n);
    for (std::size_t i = 0; i < n; ++i)
    {
        points[i] = pcp::point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        nrm[i]    = pcp::normal_t{normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    }
    std::vector<std::size_t> indices(n);
    std::iota(indices.begin(), indices.end(), std::size_t{0});
    auto const coordinate_map = [&](std::size_t const i) {
        return std::array<float, 3u>{points[i].x(), points[i].y(), points[i].z()};
    };
    auto const point_map  = [&](std::size_t const i) { return points[i]; };
    auto const index_map  = [](std::size_t const i) { return i; };
    auto const normal_map = [&](std::size_t const i) { return nrm[i]; };
    auto const transform_op = [&](std::size_t const i, pcp::normal_t const& v) { nrm[i] = v; };

    pcp::kdtree::construction_params_t params;
    params.compute_max_depth = true;
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(coordinate_map)> kdtree{
        indices.begin(), indices.end(), coordinate_map, params};
    auto const knn_map = [&](std::size_t const i) -> std::vector<std::size_t> {
        if (knn == nullptr)
            return kdtree.nearest_neighbours(i, k);
        std::vector<std::size_t> out;
        for (std::size_t j = 0; j < k; ++j)
            if (knn[i * k + j] >= 0)
                out.push_back(static_cast<std::size_t>(knn[i * k + j]));
        return out;
    };
    pcp::algorithm::propagate_normal_orientations(
        indices.begin(), indices.end(), index_map, knn_map, point_map, normal_map, transform_op);
    for (std::size_t i = 0; i < n; ++i)
        normals[3 * i] = nrm[i].nx(), normals[3 * i + 1] = nrm[i].ny(),
                    normals[3 * i + 2] = nrm[i].nz();
}
