// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product path.
//
// Flat C bridge over the UNMODIFIED reference headers for the radius-search callers
// (SURVEY.md §8f rank 3), compiled where they lie under /root/reference/include by
// oracle/Makefile into oracle/_ref/libpcp_ref_smoothing.so:
//   pcp::algorithm::bilateral_filter_points             include/pcp/algorithm/bilateral_filter.hpp:301-421
//   pcp::algorithm::wlop::detail::compute_vj / _wi      include/pcp/algorithm/wlop.hpp:28-104
//   pcp::algorithm::wlop::detail::solve_first_energy_median            :106-168
//   pcp::algorithm::wlop::detail::solve_second_energy_repulsion_force  :170-224
//
// bilateral_filter.hpp includes <Eigen/Core>; oracle/ref_shim/Eigen/Core declares the one name
// it needs to be PARSED (Eigen::Matrix).  bilateral_filter_points / compute_pi use no Eigen and
// run unmodified; bilateral_filter_normals (compute_ni, all Eigen) cannot be built here and is
// covered only by the restatement in oracle/pcp_oracle.c.
//
// pcp::algorithm::wlop::wlop itself draws its start set from std::random_device (:346-358) and
// so cannot produce a repeatable fixture; ref_wlop below runs the SAME sequence of reference
// calls as wlop.hpp:360-435 (same kd-tree parameters, same detail functions, same update) on a
// start set passed in by the caller.  Only that driver loop is ours.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <vector>

#include <pcp/algorithm/bilateral_filter.hpp>
#include <pcp/algorithm/wlop.hpp>
#include <pcp/common/normals/normal.hpp>
#include <pcp/common/points/point.hpp>
#include <pcp/kdtree/kdtree.hpp>

extern "C" {

void ref_bilateral_filter_points(const float* xyz, const float* nrm, std::size_t n, double sigmaf,
                                 double sigmag, std::size_t K, float* out)
{
    std::vector<pcp::point_t> points(n);
    std::vector<pcp::normal_t> normals(n);
    for (std::size_t i = 0; i < n; ++i)
    {
        points[i]  = pcp::point_t{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        normals[i] = pcp::normal_t{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    }
    std::vector<std::size_t> indices(n);
    std::iota(indices.begin(), indices.end(), std::size_t{0});
    auto const point_map  = [&](std::size_t const i) { return points[i]; };
    auto const normal_map = [&](std::size_t const i) { return normals[i]; };
    pcp::algorithm::bilateral::params_t params;
    params.K      = K;
    params.sigmaf = sigmaf;
    params.sigmag = sigmag;
    std::vector<pcp::point_t> filtered;
    filtered.reserve(n);
    pcp::algorithm::bilateral_filter_points(
        indices.begin(), indices.end(), std::back_inserter(filtered), point_map, normal_map,
        params);
    for (std::size_t i = 0; i < n; ++i)
        out[3 * i] = filtered[i].x(), out[3 * i + 1] = filtered[i].y(),
                out[3 * i + 2] = filtered[i].z();
}

void ref_wlop(const float* xyz, std::size_t J, const std::uint32_t* initial, std::size_t I,
              double mu_d, double h_d, std::size_t K, int uniform, float* out)
{
    namespace wd = pcp::algorithm::wlop::detail;
    using scalar = float;
    scalar const mu = static_cast<scalar>(mu_d);
    scalar const h  = static_cast<scalar>(h_d);
    std::vector<scalar> vj(J, scalar{1.0});
    std::vector<scalar> wi(I, scalar{1.0});
    std::vector<pcp::point_t> x(I), xp(I);
    std::vector<std::size_t> is(I), js(J);
    std::iota(is.begin(), is.end(), std::size_t{0});
    std::iota(js.begin(), js.end(), std::size_t{0});
    scalar const h_over_4_squared = (h * h) / scalar{4.0 * 4.0};
    auto const theta = [=](scalar const r2) { return std::exp(-r2 / h_over_4_squared); };
    auto const p_cmap = [&](std::size_t const j) {
        return std::array<scalar, 3u>{xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    };
    auto const q_cmap = [&](std::size_t const i) {
        return std::array<scalar, 3u>{x[i].x(), x[i].y(), x[i].z()};
    };
    auto const vj_map = [&](std::size_t const j) { return vj[j]; };
    auto const wi_map = [&](std::size_t const i) { return wi[i]; };
    for (std::size_t i = 0; i < I; ++i)
        x[i] = pcp::point_t{xyz[3 * initial[i]], xyz[3 * initial[i] + 1], xyz[3 * initial[i] + 2]};
    xp = x;

    pcp::kdtree::construction_params_t kp;
    kp.compute_max_depth     = true;
    kp.construction          = pcp::kdtree::construction_t::nth_element;
    kp.max_elements_per_leaf = 64u;
    pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(p_cmap)> p_kdtree{
        js.begin(), js.end(), p_cmap, kp};
    if (uniform)
        for (std::size_t j = 0; j < J; ++j)
            vj[j] = wd::compute_vj(j, h, p_kdtree, p_cmap, theta);
    for (std::size_t k = 0; k < K; ++k)
    {
        pcp::basic_linked_kdtree_t<std::size_t, 3u, decltype(q_cmap)> q_kdtree{
            is.begin(), is.end(), q_cmap, kp};
        if (uniform)
            for (std::size_t i = 0; i < I; ++i)
                wi[i] = wd::compute_wi(i, h, q_kdtree, q_cmap, theta);
        for (std::size_t ip = 0; ip < I; ++ip)
        {
            pcp::basic_point_t<scalar> const median =
                wd::solve_first_energy_median(ip, h, p_kdtree, p_cmap, q_cmap, vj_map, theta);
            pcp::common::basic_vector3d_t<scalar> const repulsion =
                wd::solve_second_energy_repulsion_force(ip, h, mu, q_kdtree, q_cmap, wi_map, theta);
            xp[ip] = pcp::point_t{
                median.x() + repulsion.x(), median.y() + repulsion.y(), median.z() + repulsion.z()};
        }
        x = xp;
    }
    for (std::size_t i = 0; i < I; ++i)
        out[3 * i] = xp[i].x(), out[3 * i + 1] = xp[i].y(), out[3 * i + 2] = xp[i].z();
}

} // extern "C"
