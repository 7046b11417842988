// End-to-end time of the reference's own call sequence through the C++ drop-in header
// (include/pcpx/pcp.hpp): std::vector<pcp::point_t> in host memory -> octree constructor ->
// estimate_normals with a gpu_knn_map -> std::vector<pcp::normal_t> in host memory.
//   g++ -std=c++17 -O2 -I include tools/cpp/e2e_dropin.cpp -L point-cloud-processing_b200/lib -lpcpx \
//       -Wl,-rpath,$PWD/point-cloud-processing_b200/lib -o tools/cpp/e2e_dropin && tools/cpp/e2e_dropin [n]
#include <pcpx/pcp.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char** argv)
{
    std::size_t const n = argc > 1 ? static_cast<std::size_t>(std::atof(argv[1])) : 10000000u;
    std::mt19937 gen(7);
    std::uniform_real_distribution<float> u(0.f, 10.f);
    std::normal_distribution<float> g(0.f, 1e-3f);
    std::vector<pcp::point_t> points;
    points.reserve(n);
    for (std::size_t i = 0; i < n; ++i)
        points.push_back({u(gen), u(gen), g(gen)});
    auto const point_map = [](pcp::point_t const& p) { return p; };
    using octree_type   = pcp::basic_linked_octree_t<pcp::point_t>;
    auto now            = [] { return std::chrono::steady_clock::now(); };
    auto ms             = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::vector<pcp::normal_t> normals(points.size());
    for (int rep = 0; rep < 5; ++rep)
    {
        auto const t0 = now();
        octree_type octree{points.begin(), points.end(), point_map};
        auto const t1 = now();
        pcp::algorithm::estimate_normals(
            std::execution::par, points.begin(), points.end(), normals.begin(), point_map,
            pcp::make_gpu_knn_map(octree, 15u, point_map),
            pcp::algorithm::default_normal_transform<pcp::point_t, pcp::normal_t>);
        auto const t2 = now();
        std::printf("rep %d: n = %zu  octree ctor %.1f ms  estimate_normals %.1f ms  total %.1f ms  "
                    "(%.1f M normals/s)  normal[0] = (%g, %g, %g)\n",
                    rep, n, ms(t0, t1), ms(t1, t2), ms(t0, t2), n / ms(t0, t2) * 1e-3,
                    normals[0].x(), normals[0].y(), normals[0].z());
    }
    return 0;
}
