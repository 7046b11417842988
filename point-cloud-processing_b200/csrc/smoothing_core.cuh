// Per-point bodies of the radius-search callers: the bilateral filter of
// points and of normals and the WLOP resampler.  Each is "gather the ball around the point from
// the index, reduce with per-neighbour weights"; the neighbour attributes (normals, density
// weights) are float arrays in the index's SORTED order so they stream with the points.
//
// All arithmetic is fp32, operation for operation as the reference writes it; the one
// unavoidable difference is the order in which neighbours are summed (the reference's order is
// its kd-tree's DFS order), so parity for these is tolerance based.
#pragma once
#include "radius_core.cuh"

namespace pcpx {

// algorithm/bilateral_filter.hpp:359-367 (`gaussian`) and :511-519 (`dgaussian`)
PCPX_HD float gaussian_w(float sigma, float r)
{
    float const s2    = sigma * sigma;
    float const r2    = r * r;
    float const power = -r2 / (2.f * s2);
    float const coeff = 1.f / (sigma * sqrtf(2.f * 3.14159265358979323846f));
    return coeff * expf(power);
}

PCPX_HD float dgaussian_w(float sigma, float r)
{
    float const s2    = sigma * sigma;
    float const s3    = sigma * s2;
    float const r2    = r * r;
    float const power = -r2 / (2.f * s2);
    float const coeff = -r / (s3 * sqrtf(2.f * 3.14159265358979323846f));
    return coeff * expf(power);
}

// bilateral::detail::compute_pi (algorithm/bilateral_filter.hpp:47-100): the point s moves to the
// weighted mean of its projections onto the neighbours' tangent planes; support = ball of radius
// 2 sigmaf, s itself included.
PCPX_HD void bilateral_point(const GridView& g, const float4* nrm_sorted, float sx, float sy,
                             float sz, float sigmaf, float sigmag, float out[3])
{
    float k = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    radius_visit(g, sx, sy, sz, 2.f * sigmaf, [&](float4 const& p, uint32_t pos) {
        float4 const n = nrm_sorted[pos];
        float const vx = p.x - sx, vy = p.y - sy, vz = p.z - sz; // sp = p - s  (:372)
        float const d  = vx * n.x + vy * n.y + vz * n.z;         // inner_product(sp, n)
        float const jx = sx + d * n.x, jy = sy + d * n.y, jz = sz + d * n.z; // s + d n
        float const fx = sx - p.x, fy = sy - p.y, fz = sz - p.z;
        float const rf = sqrtf(fx * fx + fy * fy + fz * fz);     // norm(s - p)
        float const gx = jx - sx, gy = jy - sy, gz = jz - sz;
        float const rg = sqrtf(gx * gx + gy * gy + gz * gz);     // norm(s_projected - s)
        float const w  = gaussian_w(sigmaf, rf) * gaussian_w(sigmag, rg);
        k += w;
        ax += w * jx, ay += w * jy, az += w * jz;
        return false;
    });
    out[0] = ax / k, out[1] = ay / k, out[2] = az / k;
}

PCPX_HD void normalized3(float x, float y, float z, float& ux, float& uy, float& uz)
{
    // Eigen 3.3 normalized(): the vector itself when its squared norm is not positive
    float const z2 = x * x + y * y + z * z;
    if (z2 > 0.f)
    {
        float const s = sqrtf(z2);
        ux = x / s, uy = y / s, uz = z / s;
    }
    else
        ux = x, uy = y, uz = z;
}

// bilateral::detail::compute_ni (algorithm/bilateral_filter.hpp:113-267): the normal of s is
// mapped through the Jacobian of the bilateral filter at s.  J(r, c) below is row r, column c.
PCPX_HD void bilateral_normal(const GridView& g, const float4* nrm_sorted, float sx, float sy,
                              float sz, float nsx, float nsy, float nsz, float sigmaf,
                              float sigmag, float out[3])
{
    float J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // J_pi_f_g
    float u[3] = {0, 0, 0};                    // pi_f_g
    float gk[3] = {0, 0, 0};                   // grad_k
    float k = 0.f;
    radius_visit(g, sx, sy, sz, 2.f * sigmaf, [&](float4 const& p, uint32_t pos) {
        float4 const n = nrm_sorted[pos];
        float const vx = p.x - sx, vy = p.y - sy, vz = p.z - sz;
        float const d  = vx * n.x + vy * n.y + vz * n.z;
        float const pj[3]  = {sx + d * n.x, sy + d * n.y, sz + d * n.z}; // s_projected
        float const sp[3]  = {sx - p.x, sy - p.y, sz - p.z};
        float const sps[3] = {pj[0] - sx, pj[1] - sy, pj[2] - sz};
        float const rf = sqrtf(sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2]);
        float const rg = sqrtf(sps[0] * sps[0] + sps[1] * sps[1] + sps[2] * sps[2]);
        float const wf = gaussian_w(sigmaf, rf);
        float const wg = gaussian_w(sigmag, rg);
        float const w  = wf * wg;
        k += w;
        u[0] += w * pj[0], u[1] += w * pj[1], u[2] += w * pj[2];

        float const wdf = dgaussian_w(sigmaf, rf);
        float su[3];
        normalized3(sp[0], sp[1], sp[2], su[0], su[1], su[2]);
        float const gf[3] = {su[0] * wdf, su[1] * wdf, su[2] * wdf}; // grad_f

        // "Jacobian of projection(s)" exactly as written at :213-222 (off-diagonals positive)
        float const P[9] = {1.f - n.x * n.x, n.x * n.y,       n.x * n.z,
                            n.x * n.y,       1.f - n.y * n.y, n.y * n.z,
                            n.x * n.z,       n.y * n.z,       1.f - n.z * n.z};
        float const wdg = dgaussian_w(sigmag, rg);
        float pu[3];
        normalized3(sps[0], sps[1], sps[2], pu[0], pu[1], pu[2]);
        float gg[3]; // grad_g = (sps_unit * Jpi - sps_unit) * wdg
        for (int c = 0; c < 3; ++c)
            gg[c] = ((pu[0] * P[c] + pu[1] * P[3 + c] + pu[2] * P[6 + c]) - pu[c]) * wdg;
        for (int c = 0; c < 3; ++c)
            gk[c] += gf[c] * wg + wf * gg[c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                J[3 * r + c] += (P[3 * r + c] * wf * wg + sps[r] * gf[c] * wg) +
                                sps[r] * wf * gg[c];
        return false;
    });
    float const inv = 1.f / (k * k);
    float o[3];
    for (int r = 0; r < 3; ++r)
    {
        float acc = 0.f;
        float const ns[3] = {nsx, nsy, nsz};
        for (int c = 0; c < 3; ++c)
            acc += (inv * (J[3 * r + c] * k - u[r] * gk[c])) * ns[c];
        o[r] = acc;
    }
    normalized3(o[0], o[1], o[2], out[0], out[1], out[2]);
}

// ---- WLOP (algorithm/wlop.hpp) ---------------------------------------------------------------
struct WlopParams
{
    float h;    // support radius
    float h4sq; // (h * h) / 16, the denominator of theta (:322-328)
    float mu;
};

PCPX_HD WlopParams wlop_params(float h, float mu)
{
    WlopParams w;
    w.h    = h;
    w.h4sq = (h * h) / 16.f;
    w.mu   = mu;
    return w;
}

PCPX_HD float wlop_theta(const WlopParams& w, float r2) { return expf(-r2 / w.h4sq); }

// common::are_vectors_equal(a, b, 1e-9f) (common/vector3d_queries.hpp:48-64)
PCPX_HD bool wlop_same_point(float ax, float ay, float az, float bx, float by, float bz)
{
    float const eps = 1e-9f;
    return fabsf(ax - bx) < eps && fabsf(ay - by) < eps && fabsf(az - bz) < eps;
}

// compute_vj / compute_wi (:28-104): 1 + sum of theta over the other points of the ball
PCPX_HD float wlop_density(const GridView& g, const WlopParams& w, float x, float y, float z)
{
    float v = 1.f;
    radius_visit_lazy(g, x, y, z, w.h, [&](float4 const& p, uint32_t) {
        if (!wlop_same_point(x, y, z, p.x, p.y, p.z))
            v += wlop_theta(w, sqdist_x(fsub_x(p.x, x), fsub_x(p.y, y), fsub_x(p.z, z)));
        return false;
    });
    return v;
}

// One solver step for the resampled point q (:403-432): the density-weighted local median of
// the input cloud P around q (solve_first_energy_median, :106-168) plus the repulsion of the
// other resampled points (solve_second_energy_repulsion_force, :170-224).  vj / wi are the
// density weights in the sorted order of gp / gq; nullptr = all ones (params.uniform == false).
PCPX_HD void wlop_step(const GridView& gp, const float* vj_sorted, const GridView& gq,
                       const float* wi_sorted, const WlopParams& w, float qx, float qy, float qz,
                       float out[3])
{
    float const eps = 1e-9f;
    float sum = 0.f, mx = 0.f, my = 0.f, mz = 0.f;
    radius_visit_lazy(gp, qx, qy, qz, w.h, [&](float4 const& p, uint32_t pos) {
        if (wlop_same_point(qx, qy, qz, p.x, p.y, p.z))
            return false;
        float const r2 = sqdist_x(fsub_x(p.x, qx), fsub_x(p.y, qy), fsub_x(p.z, qz));
        float const r  = sqrtf(r2);
        float const vj = vj_sorted ? vj_sorted[pos] : 1.f;
        float const alpha = fabsf(r) < eps ? 0.f : wlop_theta(w, r2) / r;
        float const coeff = fabsf(vj) < eps ? 0.f : alpha / vj;
        mx += coeff * p.x, my += coeff * p.y, mz += coeff * p.z;
        sum += coeff;
        return false;
    });
    if (fabsf(sum) < eps)
        mx = qx, my = qy, mz = qz;
    else
        mx /= sum, my /= sum, mz /= sum;

    float rsum = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;
    radius_visit_lazy(gq, qx, qy, qz, w.h, [&](float4 const& p, uint32_t pos) {
        if (wlop_same_point(p.x, p.y, p.z, qx, qy, qz))
            return false;
        float const dx = qx - p.x, dy = qy - p.y, dz = qz - p.z; // d = qip - qi
        float const r2 = sqdist_x(fsub_x(p.x, qx), fsub_x(p.y, qy), fsub_x(p.z, qz));
        float const r  = sqrtf(r2);
        float const wi = wi_sorted ? wi_sorted[pos] : 1.f;
        float const beta  = fabsf(r) < eps ? 0.f : wlop_theta(w, r2) / r;
        float const coeff = wi * beta;
        rx += coeff * dx, ry += coeff * dy, rz += coeff * dz;
        rsum += coeff;
        return false;
    });
    if (fabsf(rsum) < eps)
        rx = 0.f * rx, ry = 0.f * ry, rz = 0.f * rz;
    else
    {
        float const s = w.mu / rsum;
        rx = s * rx, ry = s * ry, rz = s * rz;
    }
    out[0] = mx + rx, out[1] = my + ry, out[2] = mz + rz;
}

} // namespace pcpx
