// Drop-in check: the reference's own test scenarios for the hot path, written against
// include/pcpx/pcp.hpp with the reference's call signatures (test/octree/octree_knn.cpp,
// test/octree/octree_range_search.cpp, test/kdtree/knn.cpp, test/common/normal_estimation.cpp,
// test/algorithm/estimate_normals.cpp, test/algorithm/average_distance_to_neighbors.cpp,
// examples/simple_example.cpp, examples/filter_point_cloud_noise_by_density.cpp,
// test/algorithm/bilateral_filter.cpp, test/algorithm/wlop.cpp).
// Exit code 0 = every scenario holds.  Needs a GPU (libpcpx has no CPU path).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <random>

#include <pcpx/pcp.hpp>
#include <pcpx/ply.hpp>
#include <sstream>

#define REQUIRE(cond)                                                                          \
    do                                                                                         \
    {                                                                                          \
        if (!(cond))                                                                           \
        {                                                                                      \
            std::fprintf(stderr, "REQUIRE failed at %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

static auto const point_map = [](pcp::point_t const& p) { return p; };

static pcp::octree_parameters_t<pcp::point_t> unit_params(float h = 1.f)
{
    pcp::octree_parameters_t<pcp::point_t> params;
    params.voxel_grid = pcp::axis_aligned_bounding_box_t<pcp::point_t>{pcp::point_t{-h, -h, -h},
                                                                       pcp::point_t{h, h, h}};
    return params;
}

static void octree_knn_scenarios()
{
    // "an octree with 1 point in each octant", k = 1
    std::vector<pcp::point_t> pts{{-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {.5f, .5f, -.5f},
                                  {-.5f, .5f, -.5f},  {-.5f, -.5f, .5f}, {.5f, -.5f, .5f},
                                  {.5f, .5f, .5f},    {-.5f, .5f, .5f}};
    pcp::linked_octree_t octree(pts.cbegin(), pts.cend(), point_map, unit_params());
    REQUIRE(octree.size() == 8u);
    auto nn = octree.nearest_neighbours(pcp::point_t{.51f, .51f, .51f}, 1u, point_map);
    REQUIRE(nn.size() == 1u && pcp::common::are_vectors_equal(nn[0], pcp::point_t{.5f, .5f, .5f}));
    nn = octree.nearest_neighbours(pcp::point_t{-.51f, -.51f, -.51f}, 1u, point_map);
    REQUIRE(nn.size() == 1u && pcp::common::are_vectors_equal(nn[0], pcp::point_t{-.5f, -.5f, -.5f}));

    // the only point equals the target -> nothing; two points, one equal -> the other one
    std::vector<pcp::point_t> one{{-.5f, -.5f, -.5f}};
    pcp::linked_octree_t o1(one.cbegin(), one.cend(), point_map, unit_params());
    REQUIRE(o1.nearest_neighbours(pcp::point_t{-.5f, -.5f, -.5f}, 1u, point_map).empty());
    std::vector<pcp::point_t> two{{-.5f, -.5f, -.5f}, {-1.f, -1.f, -1.f}};
    pcp::linked_octree_t o2(two.cbegin(), two.cend(), point_map, unit_params());
    nn = o2.nearest_neighbours(pcp::point_t{-.5f, -.5f, -.5f}, 2u, point_map);
    REQUIRE(nn.size() == 1u && pcp::common::are_vectors_equal(nn[0], pcp::point_t{-1.f, -1.f, -1.f}));

    // ordering nearest -> furthest, k = 4 and k = 3
    std::vector<pcp::point_t> ord{{-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {-.5f, .5f, -.5f},
                                  {-.5f, -.5f, .5f},  {.5f, -.5f, .5f},  {.5f, .5f, .5f},
                                  {-.5f, .5f, .5f}};
    pcp::point_t const first{.51f, .51f, -.51f}, second{.61f, .51f, -.51f},
        third{.41f, .31f, -.51f}, fourth{.71f, .21f, -.51f};
    ord.insert(ord.end(), {first, second, third, fourth});
    pcp::linked_octree_t o3(ord.cbegin(), ord.cend(), point_map, unit_params());
    pcp::point_t const reference{.5f, .5f, -.5f};
    nn = o3.nearest_neighbours(reference, 4u, point_map);
    REQUIRE(nn.size() == 4u);
    REQUIRE(pcp::common::are_vectors_equal(first, nn[0]) && pcp::common::are_vectors_equal(second, nn[1]));
    REQUIRE(pcp::common::are_vectors_equal(third, nn[2]) && pcp::common::are_vectors_equal(fourth, nn[3]));
    nn = o3.nearest_neighbours(reference, 3u, point_map);
    REQUIRE(nn.size() == 3u && pcp::common::are_vectors_equal(third, nn[2]));
    REQUIRE(o3.nearest_neighbours(reference, 0u, point_map).empty());

    // "a randomly constructed octree": k planted points near a corner are the k returned
    std::mt19937 gen(1234);
    std::uniform_real_distribution<float> c(-0.95f, 0.95f), nearc(-.99f, -.96f), farc(.96f, .99f);
    std::vector<pcp::point_t> cloud;
    for (int i = 0; i < 20000; ++i)
        cloud.push_back({c(gen), c(gen), c(gen)});
    std::size_t const k = 7;
    std::vector<pcp::point_t> planted;
    for (std::size_t i = 0; i < k; ++i)
        planted.push_back({nearc(gen), farc(gen), farc(gen)});
    cloud.insert(cloud.end(), planted.begin(), planted.end());
    pcp::linked_octree_t o4(cloud.cbegin(), cloud.cend(), point_map, unit_params(2.f));
    REQUIRE(o4.size() == cloud.size());
    nn = o4.nearest_neighbours(pcp::point_t{-1.f, 1.f, 1.f}, k, point_map);
    REQUIRE(nn.size() == k);
    for (auto const& p : nn)
        REQUIRE(std::any_of(planted.begin(), planted.end(),
                            [&](auto const& q) { return pcp::common::are_vectors_equal(p, q); }));

    // insertion test: points outside the voxel grid are not indexed
    std::vector<pcp::point_t> in{{.1f, .1f, .1f}, {-.3f, .3f, .3f}, {.9f, -.9f, -.9f}};
    auto with_out = in;
    with_out.insert(with_out.end(), {{-2.f, 0.f, 0.f}, {0.f, 2.f, 0.f}, {0.f, 0.f, -2.f}});
    pcp::linked_octree_t o5(with_out.cbegin(), with_out.cend(), point_map, unit_params());
    REQUIRE(o5.size() == in.size());
}

static void octree_range_scenarios()
{
    std::vector<pcp::point_t> pts;
    for (float z : {-1.f, 1.f})
        for (auto xy : {std::pair<float, float>{-1, -1}, {1, -1}, {1, 1}, {-1, 1}})
        {
            pts.push_back({.5f * xy.first, .5f * xy.second, .5f * z});
            pts.push_back({.4f * xy.first, .3f * xy.second, .6f * z});
        }
    pcp::linked_octree_t octree(pts.cbegin(), pts.cend(), point_map, unit_params());
    pcp::sphere_t<pcp::point_t> sphere;
    sphere.position = {0.f, 0.f, 0.f};
    sphere.radius   = 0.1f;
    REQUIRE(octree.range_search(sphere, point_map).empty());
    sphere.position = {.9f, .9f, .9f};
    sphere.radius   = 1.f;
    auto in = octree.range_search(sphere, point_map);
    REQUIRE(in.size() == 2u);
    auto has = [&](pcp::point_t q) {
        return std::count_if(in.begin(), in.end(),
                             [&](auto const& p) { return pcp::common::are_vectors_equal(p, q); }) == 1;
    };
    REQUIRE(has({.5f, .5f, .5f}) && has({.4f, .3f, .6f}));
    pcp::axis_aligned_bounding_box_t<pcp::point_t> aabb;
    aabb.min = {1.05f, 1.05f, 1.05f}, aabb.max = {2.f, 2.f, 2.f};
    REQUIRE(octree.range_search(aabb, point_map).empty());
    aabb.min = {-2.f, -2.f, -2.f}, aabb.max = {0.f, 0.f, 0.f};
    in = octree.range_search(aabb, point_map);
    REQUIRE(in.size() == 2u && has({-.5f, -.5f, -.5f}) && has({-.4f, -.3f, -.6f}));
}

// octree/linked_octree.hpp:71,179-206 (empty octree + insert), test/octree/octree_insertion.cpp;
// common/points/vertex.hpp; and a small box far from the origin (the circumscribed-sphere
// pre-filter must not lose the points in its corners to the rounding of its centre)
static void insertion_vertex_and_far_box_scenarios()
{
    pcp::linked_octree_t octree(unit_params(1.f));
    REQUIRE(octree.size() == 0u);
    std::vector<pcp::point_t> inside{{.1f, .2f, .3f}, {-.5f, .5f, .9f}, {1.f, -1.f, 1.f}},
        outside{{1.1f, 0.f, 0.f}, {0.f, -1.5f, 0.f}};
    REQUIRE(octree.insert(inside.begin(), inside.end(), point_map) == 3u);
    REQUIRE(octree.insert(outside.begin(), outside.end(), point_map) == 0u);
    REQUIRE(!octree.insert(pcp::point_t{2.f, 2.f, 2.f}, point_map));
    REQUIRE(octree.insert(pcp::point_t{.11f, .2f, .3f}, point_map));
    REQUIRE(octree.size() == 4u);
    auto nn = octree.nearest_neighbours(pcp::point_t{.1f, .2f, .3f}, 1u, point_map);
    REQUIRE(nn.size() == 1u && pcp::common::are_vectors_equal(nn[0], pcp::point_t{.11f, .2f, .3f}));

    std::vector<pcp::point_t> cloud{{0.f, 0.f, 0.f}, {1.f, 2.f, 3.f}};
    pcp::vertex_t v0(&cloud[0], 0u), v1(&cloud[1], 7u), v1b(&cloud[0], 7u);
    REQUIRE(v1.id() == 7u && v1.x() == 1.f && v1.z() == 3.f && v0 != v1 && v1 == v1b);
    v0.id(9u);
    v0.y(5.f);
    REQUIRE(v0.id() == 9u && cloud[0].y() == 5.f);

    // 9 x 9 x 9 lattice, spacing 1e-4, near x = y = z = 100; the box holds a 5 x 5 x 5 sub-lattice
    std::vector<pcp::point_t> far;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j)
            for (int l = 0; l < 9; ++l)
                far.push_back({100.f + 1e-4f * i, 100.f + 1e-4f * j, 100.f + 1e-4f * l});
    pcp::linked_octree_t foct(far.cbegin(), far.cend(), point_map);
    pcp::axis_aligned_bounding_box_t<pcp::point_t> box;
    box.min = far[(2 * 9 + 2) * 9 + 2], box.max = far[(6 * 9 + 6) * 9 + 6];
    std::size_t want = 0;
    for (auto const& p : far)
        want += box.contains(p);
    REQUIRE(want >= 125u);
    REQUIRE(foct.range_search(box, point_map).size() == want);
}

static void kdtree_scenarios()
{
    auto const coordinate_map = [](pcp::point_t const& p) {
        return std::array<float, 3u>{p.x(), p.y(), p.z()};
    };
    std::vector<pcp::point_t> ord{{-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {-.5f, .5f, -.5f},
                                  {-.5f, -.5f, .5f},  {.5f, -.5f, .5f},  {.5f, .5f, .5f},
                                  {-.5f, .5f, .5f},   {.51f, .51f, -.51f}, {.61f, .51f, -.51f},
                                  {.41f, .31f, -.51f}, {.71f, .21f, -.51f}};
    pcp::basic_linked_kdtree_t<pcp::point_t, 3u, decltype(coordinate_map)> kdtree{
        ord.begin(), ord.end(), coordinate_map};
    auto nn = kdtree.nearest_neighbours(std::array<float, 3u>{.5f, .5f, -.5f}, 4u);
    REQUIRE(nn.size() == 4u && pcp::common::are_vectors_equal(nn[0], ord[7]) &&
            pcp::common::are_vectors_equal(nn[3], ord[10]));
    nn = kdtree.nearest_neighbours(ord[7], 2u); // element overload: itself is excluded
    REQUIRE(nn.size() == 2u && pcp::common::are_vectors_equal(nn[0], ord[8]));
    pcp::sphere_a<float> ball{{.5f, .5f, -.5f}, 0.12f};
    REQUIRE(kdtree.range_search(ball).size() == 2u); // (.51,.51,-.51) and (.61,.51,-.51)

    // test/algorithm/average_distance_to_neighbors.cpp: 4 clusters of 3 collinear points
    float const d = 0.1f;
    std::vector<pcp::point_t> cl{{0, 0, 0}, {0, 0, d},  {0, 0, -d}, {1, 0, 0}, {1, d, 0},  {1, -d, 0},
                                 {0, 1, 0}, {d, 1, 0},  {-d, 1, 0}, {0, 0, 1}, {d, 0, 1},  {-d, 0, 1}};
    pcp::basic_linked_kdtree_t<pcp::point_t, 3u, decltype(coordinate_map)> kd2{
        cl.begin(), cl.end(), coordinate_map};
    float const mu = pcp::algorithm::average_distance_to_neighbors(kd2, 2u);
    REQUIRE(pcp::common::floating_point_equals(mu, (16.f / 12.f) * d));
}

static void normals_scenarios()
{
    // test/common/normal_estimation.cpp: axis cross -> +-(0,0,1), unit norm
    std::vector<pcp::point_t> cross{{0, 0, 0}, {-2, 0, 0}, {2, 0, 0}, {0, -2, 0},
                                    {0, 2, 0}, {0, 0, -1}, {0, 0, 1}};
    auto const n = pcp::estimate_normal(cross.cbegin(), cross.cend(), point_map);
    pcp::normal_t const expected{0.f, 0.f, 1.f};
    REQUIRE(pcp::common::are_vectors_equal(n, expected) ||
            pcp::common::are_vectors_equal(n, -expected));
    REQUIRE(pcp::common::floating_point_equals(pcp::common::norm(n), 1.f));

    // examples/simple_example.cpp shape: point views over a vector, octree of views, density by
    // range search, estimate_normals with a kNN callable
    std::mt19937 gen(42);
    std::normal_distribution<float> g(0.f, 1.f);
    std::vector<pcp::point_t> points;
    for (int i = 0; i < 50000; ++i)
    {
        float x = g(gen), y = g(gen), z = g(gen);
        float const r = (1.f + 0.002f * g(gen)) / std::sqrt(x * x + y * y + z * z);
        points.push_back({x * r, y * r, z * r});
    }
    std::vector<pcp::point_view_t> views;
    for (auto& p : points)
        views.push_back(pcp::point_view_t{&p});
    auto const point_view_map = [](pcp::point_view_t const& p) { return p; };
    using octree_type = pcp::basic_linked_octree_t<pcp::point_view_t>;
    octree_type octree{views.begin(), views.end(), point_view_map};
    REQUIRE(octree.size() == points.size());

    // (a) the recognised map: one fused device call for the whole range
    std::vector<pcp::normal_t> normals(points.size());
    pcp::algorithm::estimate_normals(
        std::execution::par, views.begin(), views.end(), normals.begin(), point_view_map,
        pcp::make_gpu_knn_map(octree, 15u, point_view_map),
        pcp::algorithm::default_normal_transform<pcp::point_view_t, pcp::normal_t>);
    // (b) an arbitrary user lambda, exactly as the reference's examples write it
    auto const knn = [&](pcp::point_view_t const& p) {
        return octree.nearest_neighbours(p, 15u, point_view_map);
    };
    std::size_t const sample = 200;
    std::vector<pcp::normal_t> normals_b;
    pcp::algorithm::estimate_normals(
        views.begin(), views.begin() + sample, std::back_inserter(normals_b), point_view_map, knn,
        pcp::algorithm::default_normal_transform<pcp::point_view_t, pcp::normal_t>);
    REQUIRE(normals_b.size() == sample);
    double mean_abs_dot = 0.0;
    for (std::size_t i = 0; i < points.size(); ++i)
    {
        // on a sphere the PCA normal is radial (up to sign)
        float const dot = normals[i].x() * points[i].x() + normals[i].y() * points[i].y() +
                          normals[i].z() * points[i].z();
        REQUIRE(std::abs(dot) > 0.8f);
        mean_abs_dot += std::abs(dot) / static_cast<double>(points.size());
        if (i < sample)
        {
            float const ab = normals[i].x() * normals_b[i].x() + normals[i].y() * normals_b[i].y() +
                             normals[i].z() * normals_b[i].z();
            REQUIRE(1.f - std::abs(ab) <= 1e-4f); // both routes agree (test/algorithm/estimate_normals.cpp)
        }
    }

    REQUIRE(mean_abs_dot > 0.995);

    // batched kNN == per-query kNN
    auto const batch = octree.nearest_neighbours(views.begin(), views.begin() + 50, 8u, point_view_map);
    for (std::size_t i = 0; i < 50; ++i)
    {
        auto const one = octree.nearest_neighbours(views[i], 8u, point_view_map);
        REQUIRE(batch.counts[i] == one.size());
        for (std::size_t j = 0; j < one.size(); ++j)
            REQUIRE(one[j].point() == &points[batch.indices[i * 8 + j]]);
    }

    // examples/filter_point_cloud_noise_by_density.cpp: radius = mean kNN distance, threshold 5
    std::uniform_real_distribution<float> u(-1.5f, 1.5f);
    std::size_t const n_surface = points.size();
    for (int i = 0; i < 2500; ++i)
        points.push_back({u(gen), u(gen), u(gen)});
    views.clear();
    for (auto& p : points)
        views.push_back(pcp::point_view_t{&p});
    octree_type noisy{views.begin(), views.end(), point_view_map};
    float const radius = pcp::algorithm::average_distance_to_neighbors(noisy, 15u);
    std::vector<std::uint8_t> mask;
    auto const kept = pcp::algorithm::filter_by_density(noisy, radius, 5u, &mask);
    std::size_t kept_surface = 0, kept_noise = 0;
    for (std::size_t i = 0; i < mask.size(); ++i)
        (i < n_surface ? kept_surface : kept_noise) += mask[i];
    REQUIRE(kept.size() == kept_surface + kept_noise);
    REQUIRE(kept_surface > n_surface * 95 / 100); // the surface survives
    REQUIRE(kept_noise < 2500 / 10);              // the uniform noise does not
    for (std::size_t i = 1; i < kept.size(); ++i) // stable: original relative order
        REQUIRE(kept[i - 1].point() < kept[i].point());
}

// Two GPUs behind the reference's own calls (pcp::use_devices): the index is replicated, the
// batched calls are sharded, and every result equals the one-GPU result bit for bit.
static void two_device_scenario()
{
    std::mt19937 gen(7);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    std::normal_distribution<float> g(0.f, 1.f);
    std::vector<pcp::point_t> points;
    for (int i = 0; i < 200000; ++i)
        points.push_back({u(gen), u(gen), 0.002f * g(gen)});
    for (int i = 0; i < 5000; ++i) // stray points: the warp-per-query path on both devices
        points.push_back({u(gen), u(gen), u(gen) - 0.5f});
    using octree_type = pcp::basic_linked_octree_t<pcp::point_t>;
    auto run = [&](std::vector<pcp::normal_t>& normals, pcp::knn_result_t& rows, float& mean) {
        octree_type octree{points.begin(), points.end(), point_map};
        normals.assign(points.size(), pcp::normal_t{});
        pcp::algorithm::estimate_normals(
            std::execution::par, points.begin(), points.end(), normals.begin(), point_map,
            pcp::make_gpu_knn_map(octree, 15u, point_map),
            pcp::algorithm::default_normal_transform<pcp::point_t, pcp::normal_t>);
        rows = octree.nearest_neighbours(points.begin(), points.begin() + 20000, 10u, point_map);
        mean = pcp::algorithm::average_distance_to_neighbors(octree, 15u);
    };
    std::vector<pcp::normal_t> n1, n2;
    pcp::knn_result_t r1, r2;
    float m1 = 0.f, m2 = 0.f;
    run(n1, r1, m1);
    pcp::use_devices({0, 1});
    run(n2, r2, m2);
    pcp::use_devices({});
    REQUIRE(r1.indices == r2.indices && r1.squared_distances == r2.squared_distances &&
            r1.counts == r2.counts);
    REQUIRE(m1 == m2);
    for (std::size_t i = 0; i < n1.size(); ++i)
        REQUIRE(n1[i].x() == n2[i].x() && n1[i].y() == n2[i].y() && n1[i].z() == n2[i].z());
}

// test/algorithm/bilateral_filter.cpp and test/algorithm/wlop.cpp, same calls
static void smoothing_scenarios()
{
    std::vector<pcp::point_t> points{{-0.1f, 0.f, 0.f},   {-0.075f, 0.f, 0.f}, {-0.05f, 0.f, 0.01f},
                                     {-0.025f, 0.f, 0.f}, {0.0f, 0.f, 0.f},    {0.025f, 0.f, 0.f},
                                     {0.05f, 0.f, -0.01f}, {0.075f, 0.f, 0.f}, {0.1f, 0.f, 0.f}};
    std::vector<pcp::normal_t> normals(9, pcp::normal_t{0.f, 0.f, 1.f});
    normals[2] = pcp::normal_t{-0.19611614f, 0.f, 0.98058068f};
    normals[6] = pcp::normal_t{0.19611614f, 0.f, 0.98058068f};
    std::vector<std::size_t> indices(points.size());
    std::iota(indices.begin(), indices.end(), std::size_t{0});
    auto const pmap = [&](std::size_t const i) { return points[i]; };
    auto const nmap = [&](std::size_t const i) { return normals[i]; };
    auto const cmap = [&](std::size_t const i) {
        return std::array<float, 3u>{points[i].x(), points[i].y(), points[i].z()};
    };
    pcp::kdtree::construction_params_t kp;
    kp.compute_max_depth = true;
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(cmap)> kdtree{indices.begin(),
                                                                       indices.end(), cmap, kp};
    pcp::algorithm::bilateral::params_t params;
    params.K      = 2u;
    params.sigmaf = static_cast<double>(pcp::algorithm::average_distance_to_neighbors(kdtree, 2u));
    params.sigmag = params.sigmaf / 8.;
    std::vector<pcp::point_t> filtered_points;
    pcp::algorithm::bilateral_filter_points(indices.begin(), indices.end(),
                                            std::back_inserter(filtered_points), pmap, nmap, params);
    REQUIRE(filtered_points.size() == points.size());
    REQUIRE(points[2].z() > filtered_points[2].z());
    REQUIRE(points[6].z() < filtered_points[6].z());
    std::vector<pcp::normal_t> filtered_normals;
    pcp::algorithm::bilateral_filter_normals(indices.begin(), indices.end(),
                                             std::back_inserter(filtered_normals), pmap, nmap,
                                             params);
    REQUIRE(filtered_normals.size() == indices.size());

    // WLOP: 1000 uniform points in [-10, 10]^3, I = n / 2, k = 2, h = mean 15-NN distance
    std::mt19937 gen(2024);
    std::uniform_real_distribution<float> dis(-10.f, 10.f);
    std::vector<pcp::point_t> cloud(1000);
    for (auto& p : cloud)
        p = pcp::point_t{dis(gen), dis(gen), dis(gen)};
    std::vector<std::size_t> ids(cloud.size());
    std::iota(ids.begin(), ids.end(), std::size_t{0});
    auto const cloud_map = [&](std::size_t const i) { return cloud[i]; };
    auto const cloud_cmap = [&](std::size_t const i) {
        return std::array<float, 3u>{cloud[i].x(), cloud[i].y(), cloud[i].z()};
    };
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(cloud_cmap)> cloud_tree{
        ids.begin(), ids.end(), cloud_cmap, kp};
    pcp::algorithm::wlop::params_t wp;
    wp.k       = 2u;
    wp.I       = cloud.size() / 2;
    wp.h       = static_cast<double>(pcp::algorithm::average_distance_to_neighbors(cloud_tree, 15u));
    wp.uniform = true;
    std::vector<pcp::point_t> down;
    pcp::algorithm::wlop::wlop(ids.begin(), ids.end(), std::back_inserter(down), cloud_map, wp);
    REQUIRE(down.size() == wp.I);
    for (auto const& p : down)
        REQUIRE(std::isfinite(p.x()) && std::isfinite(p.y()) && std::isfinite(p.z()));
}

// examples/normals_estimation.cpp:69-117: kd-tree -> estimate_normals -> propagate orientations
static void orientation_scenario()
{
    std::mt19937 gen(77);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::size_t const n = 20000;
    std::vector<pcp::point_t> points(n);
    for (auto& p : points)
    {
        float x = nd(gen), y = nd(gen), z = nd(gen);
        float const r = std::sqrt(x * x + y * y + z * z);
        p = pcp::point_t{x / r, y / r, z / r};
    }
    std::vector<pcp::normal_t> normals(n);
    std::vector<std::size_t> indices(n);
    std::iota(indices.begin(), indices.end(), std::size_t{0});
    auto const pmap = [&](std::size_t const i) { return points[i]; };
    auto const cmap = [&](std::size_t const i) {
        return std::array<float, 3u>{points[i].x(), points[i].y(), points[i].z()};
    };
    auto const imap = [](std::size_t const i) { return i; };
    auto nmap       = [&](std::size_t const i) { return normals[i]; };
    pcp::kdtree::construction_params_t kp;
    kp.compute_max_depth = true;
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(cmap)> kdtree{indices.begin(),
                                                                       indices.end(), cmap, kp};
    auto const knn = pcp::make_gpu_knn_map(kdtree, 12u, pmap);
    pcp::algorithm::estimate_normals(
        std::execution::par, indices.begin(), indices.end(), normals.begin(), pmap, knn,
        pcp::algorithm::default_normal_transform<std::size_t, pcp::normal_t>);
    std::size_t calls = 0;
    auto const set_normal = [&](std::size_t const i, pcp::normal_t const& v) {
        normals[i] = v;
        ++calls;
    };
    pcp::algorithm::propagate_normal_orientations(indices.begin(), indices.end(), imap, knn, pmap,
                                                  nmap, set_normal);
    REQUIRE(calls >= 1u);
    std::size_t outward = 0;
    for (std::size_t i = 0; i < n; ++i)
        outward += (normals[i].nx() * points[i].x() + normals[i].ny() * points[i].y() +
                    normals[i].nz() * points[i].z()) > 0.f;
    REQUIRE(outward > n - n / 100);

    // an arbitrary KnnMap lambda (per-element calls) gives the same orientation (2 000 points)
    std::size_t const m = 2000;
    std::vector<pcp::point_t> small(points.begin(), points.begin() + m);
    std::vector<std::size_t> ids(m);
    std::iota(ids.begin(), ids.end(), std::size_t{0});
    auto const spmap = [&](std::size_t const i) { return small[i]; };
    auto const scmap = [&](std::size_t const i) {
        return std::array<float, 3u>{small[i].x(), small[i].y(), small[i].z()};
    };
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(scmap)> stree{ids.begin(), ids.end(),
                                                                       scmap, kp};
    std::vector<pcp::normal_t> a(m), b;
    auto const sknn = pcp::make_gpu_knn_map(stree, 8u, spmap);
    pcp::algorithm::estimate_normals(
        ids.begin(), ids.end(), a.begin(), spmap, sknn,
        pcp::algorithm::default_normal_transform<std::size_t, pcp::normal_t>);
    b           = a;
    auto amap   = [&](std::size_t const i) { return a[i]; };
    auto bmap   = [&](std::size_t const i) { return b[i]; };
    auto const set_a = [&](std::size_t const i, pcp::normal_t const& v) { a[i] = v; };
    auto const set_b = [&](std::size_t const i, pcp::normal_t const& v) { b[i] = v; };
    auto const per_element = [&](std::size_t const i) { return stree.nearest_neighbours(i, 8u); };
    pcp::algorithm::propagate_normal_orientations(ids.begin(), ids.end(), imap, sknn, spmap, amap,
                                                  set_a);
    pcp::algorithm::propagate_normal_orientations(ids.begin(), ids.end(), imap, per_element, spmap,
                                                  bmap, set_b);
    for (std::size_t i = 0; i < m; ++i)
        REQUIRE(a[i].nx() == b[i].nx() && a[i].ny() == b[i].ny() && a[i].nz() == b[i].nz());
}

// examples/normals_estimation.cpp:60-66,132-140: the cloud comes from and goes back to a PLY file
static void ply_round_trip()
{
    std::vector<pcp::point_t> points{{0.f, 0.25f, -1.5f}, {1.f, 2.f, 3.f}, {-4.f, 5.5f, 6.f}};
    std::vector<pcp::normal_t> normals{{0.f, 0.f, 1.f}, {0.f, 1.f, 0.f}, {1.f, 0.f, 0.f}};
    for (auto format : {pcp::io::ply_format_t::ascii, pcp::io::ply_format_t::binary_little_endian,
                        pcp::io::ply_format_t::binary_big_endian})
    {
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        pcp::io::write_ply<pcp::point_t, pcp::normal_t>(ss, points, normals, format);
        auto [p, n] = pcp::io::read_ply<pcp::point_t, pcp::normal_t>(ss);
        REQUIRE(p.size() == points.size() && n.size() == normals.size());
        for (std::size_t i = 0; i < p.size(); ++i)
        {
            REQUIRE(pcp::common::are_vectors_equal(p[i], points[i]));
            REQUIRE(n[i].nx() == normals[i].nx() && n[i].ny() == normals[i].ny() &&
                    n[i].nz() == normals[i].nz());
        }
    }
}

// common/axis_aligned_bounding_box.hpp, common/intersections.hpp and the kd-tree box range
// (test/kdtree/kdtree_range_search.cpp:84-117)
static void geometry_helpers()
{
    std::vector<pcp::point_t> pts{{-1.f, 2.f, 0.5f}, {3.f, -4.f, 0.25f}, {0.f, 0.f, -7.f}};
    auto const box = pcp::bounding_box<std::vector<pcp::point_t>::const_iterator, pcp::point_t>(
        pts.cbegin(), pts.cend());
    REQUIRE(box.min.x() == -1.f && box.min.y() == -4.f && box.min.z() == -7.f);
    REQUIRE(box.max.x() == 3.f && box.max.y() == 2.f && box.max.z() == 0.5f);
    REQUIRE(box.contains(pcp::point_t{0.f, 0.f, 0.f}) && !box.contains(pcp::point_t{4.f, 0.f, 0.f}));
    auto const np = box.nearest_point_from(pcp::point_t{10.f, 0.f, -9.f});
    REQUIRE(np.x() == 3.f && np.y() == 0.f && np.z() == -7.f);
    pcp::sphere_t<pcp::point_t> s;
    s.position = pcp::point_t{5.f, 0.f, 0.f};
    s.radius   = 1.5f;
    REQUIRE(!pcp::intersects(box, s)); // 2 away from the box
    s.radius = 2.f;
    REQUIRE(pcp::intersects(box, s));

    std::vector<std::size_t> ids{0u, 1u, 2u};
    auto const cmap = [&](std::size_t const i) {
        return std::array<float, 3u>{pts[i].x(), pts[i].y(), pts[i].z()};
    };
    auto const kbox = pcp::kd_bounding_box<float, 3u>(ids.begin(), ids.end(), cmap);
    REQUIRE(kbox.min[0] == -1.f && kbox.max[1] == 2.f && kbox.min[2] == -7.f);
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(cmap)> kdtree{ids.begin(), ids.end(), cmap};
    pcp::kd_axis_aligned_bounding_box_t<float, 3u> q;
    q.min = {-2.f, -5.f, 0.f};
    q.max = {4.f, 3.f, 1.f};
    auto in_box = kdtree.range_search(q);
    std::sort(in_box.begin(), in_box.end());
    REQUIRE(in_box.size() == 2u && in_box[0] == 0u && in_box[1] == 1u);
    pcp::sphere_a<float> sa;
    sa.position = {0.f, 0.f, -7.f};
    sa.radius   = 0.5f;
    REQUIRE(pcp::intersects(kbox, sa));
}

int main()
{
    if (pcpx_device_count() < 1)
    {
        std::fprintf(stderr, "no CUDA device: libpcpx has no CPU path\n");
        return 2;
    }
    octree_knn_scenarios();
    octree_range_scenarios();
    insertion_vertex_and_far_box_scenarios();
    kdtree_scenarios();
    normals_scenarios();
    smoothing_scenarios();
    orientation_scenario();
    ply_round_trip();
    geometry_helpers();
    if (pcp::device_count() >= 2)
    {
        two_device_scenario();
        std::printf("dropin_test: two-device scenario holds\n");
    }
    std::printf("dropin_test: all scenarios hold\n");
    return 0;
}
