// Symmetric 3x3 eigen-decomposition for the PCA normal: the unit eigenvector of the smallest
// eigenvalue of the fp32 scatter matrix (common/normals/normal_estimation.hpp:52-73, where the
// reference calls Eigen::SelfAdjointEigenSolver<Matrix3f>).  Cyclic Jacobi on the matrix scaled
// by its largest |entry| (Eigen scales the same way before its QL iteration); a fixed number of
// sweeps so that every lane of a warp runs the same instruction stream.
#pragma once
#include "grid_core.cuh"

namespace pcpx {

struct Sym3
{
    float xx, xy, xz, yy, yz, zz;
};

#define PCPX_JACOBI_ROTATE(app, aqq, apq, arp, arq, vp0, vq0, vp1, vq1, vp2, vq2)              \
    do                                                                                         \
    {                                                                                          \
        float const apq_ = (apq);                                                              \
        if (fabsf(apq_) > 1e-30f)                                                              \
        {                                                                                      \
            float const theta = ((aqq) - (app)) / (2.f * apq_);                                \
            float const t =                                                                    \
                copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));       \
            float const c = 1.f / sqrtf(fmaf(t, t, 1.f));                                      \
            float const s = t * c;                                                             \
            (app) -= t * apq_;                                                                 \
            (aqq) += t * apq_;                                                                 \
            (apq)           = 0.f;                                                             \
            float const rp_ = (arp), rq_ = (arq);                                              \
            (arp) = c * rp_ - s * rq_;                                                         \
            (arq) = s * rp_ + c * rq_;                                                         \
            float tp, tq;                                                                      \
            tp = (vp0), tq = (vq0), (vp0) = c * tp - s * tq, (vq0) = s * tp + c * tq;          \
            tp = (vp1), tq = (vq1), (vp1) = c * tp - s * tq, (vq1) = s * tp + c * tq;          \
            tp = (vp2), tq = (vq2), (vp2) = c * tp - s * tq, (vq2) = s * tp + c * tq;          \
        }                                                                                      \
    } while (0)

// Returns the eigenvalues' relative gap (l1 - l0) / l2 via *gap when non-null.
PCPX_HD void smallest_eigenvector(Sym3 m, float& nx, float& ny, float& nz, float* gap)
{
    float scale = fmaxf(fmaxf(fabsf(m.xx), fabsf(m.yy)), fabsf(m.zz));
    scale       = fmaxf(scale, fmaxf(fmaxf(fabsf(m.xy), fabsf(m.xz)), fabsf(m.yz)));
    if (!(scale > 0.f) || !(scale < INFINITY))
    {
        // zero (or non-finite) scatter: < 2 distinct neighbours — no plane is defined
        nx = 0.f, ny = 0.f, nz = 1.f;
        if (gap)
            *gap = 0.f;
        return;
    }
    float const inv = 1.f / scale;
    float a00 = m.xx * inv, a01 = m.xy * inv, a02 = m.xz * inv;
    float a11 = m.yy * inv, a12 = m.yz * inv, a22 = m.zz * inv;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f; // V[row][col]; columns are eigenvectors
    float v10 = 0.f, v11 = 1.f, v12 = 0.f;
    float v20 = 0.f, v21 = 0.f, v22 = 1.f;
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep)
    {
        // (p,q) = (0,1): third index r = 2 -> a02 (r,p), a12 (r,q)
        PCPX_JACOBI_ROTATE(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);
        // (p,q) = (0,2): r = 1 -> a01 (p,r), a12 (r,q)
        PCPX_JACOBI_ROTATE(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);
        // (p,q) = (1,2): r = 0 -> a01 (r,p), a02 (r,q)
        PCPX_JACOBI_ROTATE(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);
    }
    // smallest eigenvalue; on exact ties the reference's cascade of ifs lets the later column
    // of the ascending order win (:60-73) — any member of a tied eigenspace is equally valid.
    float l0 = a00, l1 = a11, l2 = a22;
    float ex = v00, ey = v10, ez = v20, lmin = l0;
    if (l1 < lmin)
        ex = v01, ey = v11, ez = v21, lmin = l1;
    if (l2 < lmin)
        ex = v02, ey = v12, ez = v22, lmin = l2;
    float const nrm = 1.f / sqrtf(ex * ex + ey * ey + ez * ez);
    nx = ex * nrm, ny = ey * nrm, nz = ez * nrm;
    if (gap)
    {
        float const lmax = fmaxf(l0, fmaxf(l1, l2));
        float const lmid = (l0 + l1 + l2) - lmin - lmax;
        *gap             = (lmid - lmin) / fmaxf(lmax, 1e-30f);
    }
}

// Closed form of the same result for the fused kernels (the Jacobi sweeps above cost ~1000
// instructions per query and 300 of code): the smallest eigenvalue of the scaled matrix by the
// trigonometric solution of the characteristic cubic, polished by two Newton steps taken from
// just left of it (p(l) = l^3 - c2 l^2 + c1 l - c0 is negative, increasing and concave left of
// its smallest root, so the iterates rise monotonically to it), then the eigenvector as the
// largest cross product of two rows of (A - l0 I).  An error d in l0 tilts the vector by
// ~ d / (l1 - l0); with d ~ 1e-6 (fp32 cancellation in the determinant of the scaled matrix) that
// is far inside the 1e-4 |cos| tolerance wherever the normal is defined at all (relative
// eigengap > 1e-3).  A (near-)double smallest eigenvalue makes every cross product vanish; such
// matrices, and anything non-finite, take the Jacobi path.
PCPX_HD void smallest_eigenvector_fast(Sym3 m, float& nx, float& ny, float& nz)
{
    float scale = fmaxf(fmaxf(fabsf(m.xx), fabsf(m.yy)), fabsf(m.zz));
    scale       = fmaxf(scale, fmaxf(fmaxf(fabsf(m.xy), fabsf(m.xz)), fabsf(m.yz)));
    if (!(scale > 0.f) || !(scale < INFINITY))
    {
        nx = 0.f, ny = 0.f, nz = 1.f;
        return;
    }
    float const inv = 1.f / scale;
    float const a00 = m.xx * inv, a01 = m.xy * inv, a02 = m.xz * inv;
    float const a11 = m.yy * inv, a12 = m.yz * inv, a22 = m.zz * inv;
    float const c2 = a00 + a11 + a22;
    float const c1 = (a00 * a11 - a01 * a01) + (a00 * a22 - a02 * a02) + (a11 * a22 - a12 * a12);
    float const c0 = a00 * (a11 * a22 - a12 * a12) - a01 * (a01 * a22 - a12 * a02) +
                     a02 * (a01 * a12 - a11 * a02);
    float const q   = c2 * (1.f / 3.f);
    float const b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    float const p2  = b00 * b00 + b11 * b11 + b22 * b22 + 2.f * (a01 * a01 + a02 * a02 + a12 * a12);
    float const p   = sqrtf(p2 * (1.f / 6.f));
    float const ip  = p > 1e-20f ? 1.f / p : 0.f;
    float const detb = b00 * (b11 * b22 - a12 * a12) - a01 * (a01 * b22 - a12 * a02) +
                       a02 * (a01 * a12 - b11 * a02);
    float r = 0.5f * detb * ip * ip * ip;
    r       = fminf(1.f, fmaxf(-1.f, r));
    float const phi = acosf(r) * (1.f / 3.f);
#ifdef __CUDA_ARCH__
    float const cs = __cosf(phi + 2.0943951f);
#else
    float const cs = cosf(phi + 2.0943951f);
#endif
    float l = q + 2.f * p * cs - 2e-5f; // just left of the smallest root
#pragma unroll
    for (int it = 0; it < 2; ++it)
    {
        float const pl = ((l - c2) * l + c1) * l - c0;
        float const dp = (3.f * l - 2.f * c2) * l + c1;
        float const st = dp > 1e-12f ? pl / dp : 0.f;
        l              = l - fminf(st, 0.f); // p <= 0 left of the root: steps only ever go right
    }
    float ex = 0.f, ey = 0.f, ez = 0.f, best = 0.f;
    // Two rounds: the root of the cubic is ill-conditioned when l0 and l1 are close (an fp32
    // error d in p(l) moves the root by d / ((l1 - l0)(l2 - l0))), so the first vector may be
    // tilted by an angle t; its Rayleigh quotient v^T A v is off by only t^2 (l1 - l0), and the
    // vector recomputed for THAT value is off by t^2 — the gap cancels.
#pragma unroll
    for (int round = 0; round < 2; ++round)
    {
        if (round == 1)
        {
            float const nrm2 = 1.f / best;
            float const wx = a00 * ex + a01 * ey + a02 * ez, wy = a01 * ex + a11 * ey + a12 * ez,
                        wz = a02 * ex + a12 * ey + a22 * ez;
            l = (ex * wx + ey * wy + ez * wz) * nrm2;
        }
        float const r0x = a00 - l, r1y = a11 - l, r2z = a22 - l;
        // rows r0 = (r0x, a01, a02), r1 = (a01, r1y, a12), r2 = (a02, a12, r2z)
        float const ax = a01 * a12 - a02 * r1y, ay = a02 * a01 - r0x * a12, az = r0x * r1y - a01 * a01; // r0 x r1
        float const bx = a01 * r2z - a02 * a12, by = a02 * a02 - r0x * r2z, bz = r0x * a12 - a01 * a02; // r0 x r2
        float const cx = r1y * r2z - a12 * a12, cy = a12 * a02 - a01 * r2z, cz = a01 * a12 - r1y * a02; // r1 x r2
        float const na = ax * ax + ay * ay + az * az, nb = bx * bx + by * by + bz * bz,
                    nc = cx * cx + cy * cy + cz * cz;
        ex = ax, ey = ay, ez = az, best = na;
        if (nb > best)
            ex = bx, ey = by, ez = bz, best = nb;
        if (nc > best)
            ex = cx, ey = cy, ez = cz, best = nc;
        if (!(best > 1e-7f)) // (near-)double eigenvalue, or NaN
        {
            smallest_eigenvector(m, nx, ny, nz, nullptr);
            return;
        }
    }
#ifdef __CUDA_ARCH__
    float const nrm = rsqrtf(best);
#else
    float const nrm = 1.f / sqrtf(best);
#endif
    nx = ex * nrm, ny = ey * nrm, nz = ez * nrm;
}

} // namespace pcpx
