"""Synthetic clouds of the BASELINE configs.  Deterministic in (name, n, seed):
numpy's PCG64 streams, fp32 output.  The same buffers feed the oracle, the CPU baseline and the
GPU path."""
import numpy as np


def noisy_sphere(n, seed=42, sigma=0.005):
    """direction = normalised N(0,1)^3, radius = 1 + sigma * N(0,1)"""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = 1.0 + sigma * rng.standard_normal((n, 1))
    return np.ascontiguousarray((d * r).astype(np.float32))


def plane_extent(n, density=1e5):
    """side L of the square so that the in-plane density is `density` points per unit area
    (canonical: n = 1e7 -> L = 10, r = 0.01 -> ~31 neighbours)"""
    return float(np.sqrt(n / density))


def noisy_plane(n, seed=7, extent=None, sigma=1e-3):
    """x, y ~ U[0, L), z = sigma * N(0,1)"""
    rng = np.random.default_rng(seed)
    L = plane_extent(n) if extent is None else extent
    out = np.empty((n, 3), np.float32)
    out[:, 0] = rng.uniform(0.0, L, n)
    out[:, 1] = rng.uniform(0.0, L, n)
    out[:, 2] = sigma * rng.standard_normal(n)
    return out


def noise_mix(n, seed=11, noise_fraction=0.05, extent=None):
    """(1 - f) n noisy-plane points + f n points uniform in the inflated bbox cube, shuffled"""
    rng = np.random.default_rng(seed)
    n_noise = int(round(n * noise_fraction))
    n_surf = n - n_noise
    L = plane_extent(n_surf) if extent is None else extent
    surf = noisy_plane(n_surf, seed=seed + 1, extent=L)
    lo = np.array([0.0, 0.0, -0.05 * L]) - 0.02 * L
    hi = np.array([L, L, 0.05 * L]) + 0.02 * L
    noise = rng.uniform(lo, hi, (n_noise, 3)).astype(np.float32)
    pts = np.concatenate([surf, noise], 0)
    rng.shuffle(pts, axis=0)
    return np.ascontiguousarray(pts)


def scan(n, seed=13):
    """height field z = 0.05 sin(2 pi x / 5) cos(2 pi y / 5) + 1e-3 N(0,1) over [0, 30)^2 (density
    scaled to n) plus a unit noisy sphere (5 % of the points) resting on it"""
    rng = np.random.default_rng(seed)
    n_sph = n // 20
    n_hf = n - n_sph
    side = 30.0 * np.sqrt(n / 1e8) if n < 1e8 else 30.0
    x = rng.uniform(0.0, side, n_hf)
    y = rng.uniform(0.0, side, n_hf)
    z = 0.05 * np.sin(2 * np.pi * x / 5.0) * np.cos(2 * np.pi * y / 5.0) \
        + 1e-3 * rng.standard_normal(n_hf)
    hf = np.stack([x, y, z], 1)
    sph = noisy_sphere(n_sph, seed=seed + 1).astype(np.float64) * min(1.0, side / 4.0)
    sph += np.array([side / 2, side / 2, min(1.0, side / 4.0) + 0.05])
    pts = np.concatenate([hf, sph], 0).astype(np.float32)
    rng.shuffle(pts, axis=0)
    return np.ascontiguousarray(pts)


def uniform_cube(n, seed=3, half=100.0):
    """the reference benchmark's shape: uniform in [-half, half]^3
    (benchmark/spatial_data_structures_benchmark.cpp:381-494)"""
    rng = np.random.default_rng(seed)
    return rng.uniform(-half, half, (n, 3)).astype(np.float32)


def scan_slab(n_total, lo_frac, hi_frac, seed=13):
    """The part of a `scan`-shaped cloud of n_total points whose x lies in
    [lo_frac, hi_frac) * side, generated without materialising the whole cloud (one rank's share
    of configs[3]): the height field is sampled directly inside the slab, the sphere's points
    are drawn in full (5 % of the cloud) and cut.  Same distribution as `scan`, its own sample."""
    rng = np.random.default_rng([seed, int(lo_frac * 1e6), int(hi_frac * 1e6)])
    n_sph = n_total // 20
    n_hf = n_total - n_sph
    side = 30.0 * np.sqrt(n_total / 1e8) if n_total < 1e8 else 30.0
    lo, hi = lo_frac * side, hi_frac * side
    m = int(round(n_hf * (hi_frac - lo_frac)))
    x = rng.uniform(lo, hi, m)
    y = rng.uniform(0.0, side, m)
    z = 0.05 * np.sin(2 * np.pi * x / 5.0) * np.cos(2 * np.pi * y / 5.0) \
        + 1e-3 * rng.standard_normal(m)
    hf = np.stack([x, y, z], 1).astype(np.float32)
    sph = noisy_sphere(n_sph, seed=seed + 1).astype(np.float64) * min(1.0, side / 4.0)
    sph += np.array([side / 2, side / 2, min(1.0, side / 4.0) + 0.05])
    sph = sph[(sph[:, 0] >= lo) & (sph[:, 0] < hi)].astype(np.float32)
    pts = np.concatenate([hf, sph], 0)
    rng.shuffle(pts, axis=0)
    return np.ascontiguousarray(pts), side
