"""Synthetic velocity models / sources of the benchmark configurations (BASELINE.json `configs`,
SURVEY.md section 8d).  numpy only; deterministic."""
from __future__ import annotations

import math

import numpy as np


def _gauss_kernel(sigma):
    r = int(math.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def smooth_random_field(shape, sigma, seed):
    """Uniform random field, separable Gaussian smoothing (edge-replicated), rescaled to [0,1]."""
    rng = np.random.default_rng(seed)
    u = rng.random(shape, dtype=np.float64)
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    for ax in range(len(shape)):
        pad = [(0, 0)] * len(shape)
        pad[ax] = (r, r)
        up = np.pad(u, pad, mode="edge")
        u = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), ax, up)
    u -= u.min()
    u /= u.max()
    return u


def point_sources_top_grid(nodes, g1, g2=None):
    """1-based subscripts of a g1 (x g2) grid of sources on the top plane (last dim index 1)."""
    nodes = [int(v) for v in nodes]
    if len(nodes) == 2:
        xs = np.linspace(1, nodes[0], g1 + 2)[1:-1].round().astype(int)
        return [[int(x), 1] for x in xs]
    g2 = g2 or g1
    xs = np.linspace(1, nodes[0], g1 + 2)[1:-1].round().astype(int)
    ys = np.linspace(1, nodes[1], g2 + 2)[1:-1].round().astype(int)
    return [[int(x), int(y), 1] for y in ys for x in xs]


def config1():
    """test/ShiftedLaplacianTest.jl:14-45: 257x129 nodes, v = 1.5 km/s, f = 2.5 Hz."""
    v = 1.5 * np.ones((257, 129))
    return dict(domain=[0.0, 13.5, 0.0, 4.2], n_cells=[256, 128], m=1.0 / v**2, f=2.5, pad=[16, 16], gamma0_frac=0.01,
                shift=0.02, levels=2, cycle="W", relax_param=0.75)


def config2(vp_ms):
    """2-D SEG salt model (examples/SEGmodel2Dsalt.dat, 128 x 256 m/s): transpose, km/s, edge-replicate to
    257 x 129 nodes."""
    v = np.asarray(vp_ms, dtype=np.float64).T * 1e-3  # 256 x 128
    v = np.pad(v, ((0, 1), (0, 1)), mode="edge")
    return dict(domain=[0.0, 13.5, 0.0, 4.2], n_cells=[256, 128], m=1.0 / v**2, pad=[16, 16], gamma0_frac=0.01, shift=0.2,
                levels=3, cycle="V", relax_param=0.75)


def config3(n=129):
    """3-D layered model with attenuation: nodes n^3, h = 0.1 km, v(z) = 1.5 + 0.5 floor(8 z / L)."""
    L = 0.1 * (n - 1)
    z = np.linspace(0.0, L, n)
    layer = np.minimum(np.floor(8 * z / L), 7.0)
    v = np.broadcast_to(1.5 + 0.5 * layer, (n, n, n)).copy()
    att = np.broadcast_to(1.0 + layer / 8.0, (n, n, n)).copy()
    return dict(domain=[0.0, L, 0.0, L, 0.0, L], n_cells=[n - 1] * 3, m=1.0 / v**2, att_profile=att, pad=[12, 12, 12],
                gamma0_frac=0.02, shift=0.2, levels=3, cycle="V", relax_param=0.8)


def config4(n=257, sigma=8.0, seed=1234, pad=16):
    """3-D random-smooth model: v = 1.5 + 3 U, U Gaussian-smoothed uniform noise rescaled to [0,1]."""
    L = 0.1 * (n - 1)
    U = smooth_random_field((n, n, n), sigma, seed)
    v = 1.5 + 3.0 * U
    return dict(domain=[0.0, L, 0.0, L, 0.0, L], n_cells=[n - 1] * 3, m=1.0 / v**2, pad=[pad] * 3, gamma0_frac=0.01,
                shift=0.2, levels=3, cycle="V", relax_param=0.8)


# ---- config 5: the grid that is split into slabs over several GPUs -------------------------------------------
def smooth_random_field_fast(shape, sigma, seed):
    """Same field as smooth_random_field (same random stream, same truncated Gaussian, edge replication), evaluated
    with scipy.ndimage so that 513^3 takes seconds instead of hours."""
    from scipy.ndimage import correlate1d

    rng = np.random.default_rng(seed)
    u = rng.random(shape, dtype=np.float64)
    k = _gauss_kernel(sigma)
    for ax in range(len(shape)):
        u = correlate1d(u, k, axis=ax, mode="nearest")
    u -= u.min()
    u /= u.max()
    return u


def abl3d_planes(nodes, neumann_on_top, pad, amp, k0, k1):
    """getABL (src/GetHelmholtz.jl:164-218) restricted to the planes k0 <= k < k1 of the last dimension: the profile
    is separable, so a process that holds one slab never forms the whole-grid array.  Matches hh_get_abl."""
    nodes = [int(v) for v in nodes]
    g = []
    for d in range(3):
        nd, p = nodes[d], int(pad[d])
        x0, x1 = (-1.0, 1.0) if d < 2 else (0.0, 1.0)
        t = np.arange(nd, dtype=np.float64) / (nd - 1)
        x = (1.0 - t) * x0 + t * x1
        gd = np.zeros(nd)
        if not (d == 2 and neumann_on_top):
            gd[:p] += (x[:p] - x[p - 1]) ** 2
        gd[nd - p:] += (x[nd - p:] - x[nd - p]) ** 2
        gd /= (gd.max() + 1e-5)
        g.append(gd)
    v = (g[0][:, None, None] + g[1][None, :, None] + g[2][None, None, k0:k1]) * amp
    return np.minimum(v, amp)


def config5(n=513, sigma=16.0, seed=1234, pad=24, planes=None):
    """Config 4's recipe on the 513^3 grid (sigma and pad doubled).  `planes` = (k0, k1): return m for those planes
    of the last dimension only (what one slab's process needs); max_m is the whole-grid maximum (for omega_max)."""
    L = 0.05 * (n - 1)  # 25.6 km at 513 nodes: h = 0.05 km, so 10 points per wavelength doubles config 4's frequency
    U = smooth_random_field_fast((n, n, n), sigma, seed)
    v_min = 1.5 + 3.0 * float(U.min())
    k0, k1 = planes if planes is not None else (0, n)
    v = 1.5 + 3.0 * U[:, :, k0:k1]
    del U
    return dict(domain=[0.0, L, 0.0, L, 0.0, L], n_cells=[n - 1] * 3, m=np.asfortranarray(1.0 / v**2), max_m=1.0 / v_min**2,
                pad=[pad] * 3, gamma0_frac=0.01, shift=0.2, levels=3, cycle="W", relax_param=0.8, planes=(k0, k1))
