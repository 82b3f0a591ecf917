"""Host-side synthetic lattice fields in MILC's host layout (numpy).

Everything here produces arrays exactly as a MILC application would hand them to the
solver boundary (reference conventions):

* site order  ``i = node_index(x,y,z,t)``: lexicographic ``lex = x + nx*(y + ny*(z + nz*t))``,
  even sites first (``lex/2``), then odd sites (``(lex+V)/2``) --
  ``generic/layout_hyper_prime.c:509-520``;
* color vectors ``su3_vector[V]`` -> float array ``(V, 3, 2)`` (re, im) --
  ``include/milc_datatypes.h:49,56``;
* links ``fat[4*i+dir]``, ``lng[4*i+dir]`` -> ``(V, 4, 3, 3, 2)``, row-major ``e[row][col]`` --
  ``generic_ks/dslash_fn.c:453,458``, ``include/milc_datatypes.h:48,55``.

The synthetic "HISQ-like" links follow SURVEY.md section 8(d) config 2: thin links U are
random SU(3); fat links are KS-phased, non-unitary (U plus a small general complex
perturbation); long links are KS-phased ``c3 * U(x)U(x+mu)U(x+2mu)`` (so they are a scaled
U(3) matrix, which is what makes reconstruct-13 compression legitimate); antiperiodic
time boundary signs are folded in as MILC does (``generic_ks/fermion_links_fn_twist_milc.c:137-141``).
"""
import numpy as np

EVEN, ODD, EVENANDODD = 2, 1, 3


def volume(dims):
    nx, ny, nz, nt = dims
    return nx * ny * nz * nt


def lex_to_milc(dims):
    """Permutation ``idx[lex] = node_index`` for the whole lattice."""
    nx, ny, nz, nt = dims
    V = volume(dims)
    lex = np.arange(V, dtype=np.int64)
    x = lex % nx
    y = (lex // nx) % ny
    z = (lex // (nx * ny)) % nz
    t = lex // (nx * ny * nz)
    par = (x + y + z + t) & 1
    return np.where(par == 0, lex // 2, (lex + V) // 2)


def coords_lex(dims):
    nx, ny, nz, nt = dims
    V = volume(dims)
    lex = np.arange(V, dtype=np.int64)
    return lex % nx, (lex // nx) % ny, (lex // (nx * ny)) % nz, lex // (nx * ny * nz)


def random_su3(rng, n):
    """n Haar-distributed SU(3) matrices, complex128 (n,3,3): Gram-Schmidt on two Gaussian
    rows, third row = conjugate cross product (so det = 1 exactly)."""
    a = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    b = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    a /= np.sqrt(np.sum(np.abs(a) ** 2, axis=1))[:, None]
    b -= np.sum(np.conj(a) * b, axis=1)[:, None] * a
    b /= np.sqrt(np.sum(np.abs(b) ** 2, axis=1))[:, None]
    c = np.conj(np.cross(a, b))
    return np.stack([a, b, c], axis=1)


def _c2r(a):
    return np.ascontiguousarray(np.stack([a.real, a.imag], axis=-1))


def ks_phases(dims):
    """eta_mu(x) in lex order, shape (V,4): eta_x=(-1)^t, eta_y=(-1)^(t+x),
    eta_z=(-1)^(t+x+y), eta_t=1  (generic_ks/rephase.c:83-115)."""
    x, y, z, t = coords_lex(dims)
    eta = np.ones((volume(dims), 4))
    eta[:, 0] = 1 - 2 * (t & 1)
    eta[:, 1] = 1 - 2 * ((t + x) & 1)
    eta[:, 2] = 1 - 2 * ((t + x + y) & 1)
    return eta


def make_links(dims, seed=1234, fat_noise=0.05, c3=-1.0 / 24.0, eps_naik=0.0, dtype=np.float64):
    """Synthetic HISQ-like (fat, lng) in MILC host layout, shape (V,4,3,3,2)."""
    nx, ny, nz, nt = dims
    V = volume(dims)
    rng = np.random.default_rng(seed)
    U = random_su3(rng, 4 * V).reshape(nt, nz, ny, nx, 4, 3, 3)  # lex order, x fastest
    eta = ks_phases(dims).reshape(nt, nz, ny, nx, 4)
    fat = np.empty_like(U)
    lng = np.empty_like(U)
    axis_of = {0: 3, 1: 2, 2: 1, 3: 0}
    for mu in range(4):
        ax = axis_of[mu]
        u0 = U[..., mu, :, :]
        u1 = np.roll(u0, -1, axis=ax)
        u2 = np.roll(u0, -2, axis=ax)
        noise = fat_noise * (rng.standard_normal(u0.shape) + 1j * rng.standard_normal(u0.shape))
        e = eta[..., mu][..., None, None]
        fat[..., mu, :, :] = e * (u0 + noise)
        lng[..., mu, :, :] = e * (c3 * (1.0 + eps_naik)) * (u0 @ u1 @ u2)
    # antiperiodic time boundary: fat t-links on slice nt-1, long t-links on nt-3..nt-1
    fat[nt - 1, ..., 3, :, :] *= -1.0
    lng[nt - 3:, ..., 3, :, :] *= -1.0
    perm = lex_to_milc(dims)
    fat_m = np.empty((V, 4, 3, 3), dtype=np.complex128)
    lng_m = np.empty((V, 4, 3, 3), dtype=np.complex128)
    fat_m[perm] = fat.reshape(V, 4, 3, 3)
    lng_m[perm] = lng.reshape(V, 4, 3, 3)
    return _c2r(fat_m).astype(dtype), _c2r(lng_m).astype(dtype)


def make_thin_links(dims, seed=4321, spread=0.4, dtype=np.float64):
    """Thin SU(3) gauge links with KS phases and the antiperiodic time boundary folded in
    (what a MILC application hands to the fermion-link construction after rephase(ON),
    generic_ks/rephase.c:83-115, phases_in = 1), MILC host layout (V,4,3,3,2).
    U = exp(i * spread * H), H Gaussian Hermitian traceless: spread ~ 0.4 resembles a
    thermalised configuration, spread >= 3 is close to Haar-random (strong coupling)."""
    nx, ny, nz, nt = dims
    V = volume(dims)
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((4 * V, 3, 3)) + 1j * rng.standard_normal((4 * V, 3, 3))
    h = 0.5 * (a + np.conj(np.swapaxes(a, 1, 2)))
    h -= np.trace(h, axis1=1, axis2=2)[:, None, None] * np.eye(3) / 3.0
    w, v = np.linalg.eigh(h)
    U = (v * np.exp(1j * spread * w)[:, None, :]) @ np.conj(np.swapaxes(v, 1, 2))
    U = U.reshape(nt, nz, ny, nx, 4, 3, 3)
    eta = ks_phases(dims).reshape(nt, nz, ny, nx, 4)
    U = U * eta[..., None, None]
    U[nt - 1, ..., 3, :, :] *= -1.0
    perm = lex_to_milc(dims)
    out = np.empty((V, 4, 3, 3), dtype=np.complex128)
    out[perm] = U.reshape(V, 4, 3, 3)
    return _c2r(out).astype(dtype)


def make_source(dims, seed=5678, parity=EVEN, dtype=np.float64):
    """Gaussian random colour vector on `parity` sites (zero elsewhere), shape (V,3,2)."""
    V = volume(dims)
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((V, 3, 2))
    vh = V // 2
    if parity == EVEN:
        v[vh:] = 0
    elif parity == ODD:
        v[:vh] = 0
    return np.ascontiguousarray(v.astype(dtype))


def point_source(dims, coords=(0, 0, 0, 0), color=0, dtype=np.float64):
    nx, ny, nz, nt = dims
    V = volume(dims)
    x, y, z, t = coords
    lex = x + nx * (y + ny * (z + nz * t))
    i = lex // 2 if ((x + y + z + t) & 1) == 0 else (lex + V) // 2
    v = np.zeros((V, 3, 2), dtype=dtype)
    v[i, color, 0] = 1.0
    return v


def parity_slice(dims, parity):
    V = volume(dims)
    vh = V // 2
    if parity == EVEN:
        return slice(0, vh)
    if parity == ODD:
        return slice(vh, V)
    return slice(0, V)


def rhmc_offsets(n=12, mass=0.05):
    """A representative RHMC pole ladder: 4m^2 plus a geometric ladder."""
    base = 4.0 * mass * mass
    ladder = np.concatenate([[0.0], np.geomspace(1e-4, 2.0, n - 1)])
    return base + ladder
