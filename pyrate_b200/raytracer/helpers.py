"""Pilot-bundle generators (host API mirror of reference raytracer/helpers.py:
choose_nearest :45-75, build_pilotbundle :78-203, build_pilotbundle_complex :206-316).

A pilot bundle is a chief ray (column 0) plus a small cloud of neighbours in position
and direction; it is traced through `seqtrace` like any other bundle (on the device)
and the linear map between its clouds at two surfaces is the XYUV transfer matrix
(`OpticalElement.calculateXYUV`, xyuv.py).  Generation itself is O(100..2000) rays of
host arithmetic.

The reference obtains k and E of the start medium from a generalised 3x3 eigen-solve
PER RAY (`MaxwellMaterial.sortKnormUnitEField`, material/material.py:155-184 ->
`calcKnormEigenvectorsDirection` :455-497).  For the homogeneous isotropic start
media this package supports the four solutions are known in closed form:
k = +-n e / sqrt(e.e) (each twice, E anywhere in the plane E.e = 0), sorted by S.normal
like the reference does (:177-182); the E the reference gets out of LAPACK is an
arbitrary vector of that plane, ours is a deterministic one.
"""
import numpy as np

from .globalconstants import standard_wavelength
from .ray import RayBundle


def rodrigues(angle, axis):
    """Rotation matrix about the unit vector `axis` (reference helpers_math.py:69-83)."""
    mat = np.array([[0., -axis[2], axis[1]],
                    [axis[2], 0., -axis[0]],
                    [-axis[1], axis[0], 0.]])
    return np.eye(3) + np.sin(angle) * mat + (1. - np.cos(angle)) * np.dot(mat, mat)


def choose_nearest(kvec, kvecs_new, returnindex=False):
    """Per ray, the solution of `kvecs_new` (4, 3, N) nearest to `kvec` (3, N) that is
    not (nearly) `kvec` itself -- distance^2 > 1e-3, reference :45-75 (the non-
    conjugated product `dot(conj(v), v)` is the Hermitian norm)."""
    tol = 1e-3
    kvec = np.asarray(kvec)
    kvecs_new = np.asarray(kvecs_new)
    res = np.zeros_like(kvec)
    choosing_index = 0
    if kvecs_new.shape[1] == kvec.shape[0] and kvecs_new.shape[2] == kvec.shape[1]:
        diff = kvecs_new - kvec[None]
        d2 = np.real(np.sum(np.conj(diff) * diff, axis=1))           # (4, N)
        d2 = np.where(d2 > tol, d2, np.inf)
        d2 = np.where(d2 < 1e10, d2, np.inf)
        idx = np.argmin(d2, axis=0)                # first minimum, like the loop
        idx = np.where(np.isfinite(np.min(d2, axis=0)), idx, 0)
        res = kvecs_new[idx, :, np.arange(kvec.shape[1])].T.copy()
        if idx.size:
            choosing_index = int(idx[-1])
    if returnindex:
        return (choosing_index, res)
    return res


def _lspace(num_pts_dir):
    """(0, -1 .. <0, 1 .. >0): the reference's symmetric sampling (:118-129)."""
    n = num_pts_dir
    if n % 2 == 1:
        n -= 1
    return np.hstack((0, np.linspace(-1, 0, n // 2, endpoint=False),
                      np.linspace(1, 0, n // 2, endpoint=False)))


def _isotropic_index(mat, wave):
    names = {c.__name__ for c in type(mat).__mro__}
    if "IsotropicMaterial" not in names or "IsotropicGrinMaterial" in names:
        raise NotImplementedError(
            "pilot bundles start in a homogeneous isotropic medium (got %s)" %
            type(mat).__name__)
    return complex(mat.get_optical_index(None, wave))


def _perpendicular(e):
    """Unit E (Hermitian norm) with the bilinear product E.e = 0 per column."""
    e = np.asarray(e, dtype=complex)
    axis = np.zeros_like(e)
    axis[np.argmin(np.abs(e), axis=0), np.arange(e.shape[1])] = 1.0
    t = axis - np.sum(axis * e, axis=0) / np.sum(e * e, axis=0) * e
    return t / np.sqrt(np.real(np.sum(np.conj(t) * t, axis=0)))


def sorted_unit_modes_isotropic(n_index, kd, normal):
    """Closed form of `sortKnormUnitEField` (material/material.py:155-184) for a
    homogeneous isotropic medium: (k4, E4) of shape (4, 3, N), ascending S.normal
    (rows 0, 1: the backward pair, rows 2, 3: the forward pair)."""
    kd = np.asarray(kd, dtype=complex)
    ee = np.sum(kd * kd, axis=0)
    kplus = np.sqrt(n_index ** 2 / ee) * kd          # :490 eigenvalues +sqrt(w), -sqrt(w)
    e1 = _perpendicular(kd)
    # second polarisation: E2 = k x conj(E1) direction, again with E2.k = 0
    e2 = np.cross(kd, e1, axis=0)
    e2 = e2 / np.sqrt(np.real(np.sum(np.conj(e2) * e2, axis=0)))
    # S = Re(|E|^2 k - (E.k) conj(E)) = Re(k) for unit E with E.k = 0 (:214-223)
    sn = np.sum(np.real(kplus) * np.asarray(normal, dtype=float), axis=0)
    fwd = sn >= 0
    k4 = np.empty((4,) + kd.shape, dtype=complex)
    k_back = np.where(fwd, -kplus, kplus)
    k_fwd = np.where(fwd, kplus, -kplus)
    (k4[0], k4[1], k4[2], k4[3]) = (k_back, k_back, k_fwd, k_fwd)
    e4 = np.stack((e1, e2, e1, e2))
    return (k4, e4)


def _finish(surfobj, mat, xlocobj, kconek, lck, wave):
    lcobj = surfobj.rootcoordinatesystem
    # (the start medium is homogeneous: no dependence on the start points, which the
    # reference hands to the eigen-solver as well, :176-183)
    kconemat = mat.lc.returnOtherToActualDirections(kconek, lck)
    xlocsurf = surfobj.shape.lc.returnOtherToActualPoints(xlocobj, lcobj)
    surfnormalmat = mat.lc.returnOtherToActualDirections(
        np.asarray(surfobj.shape.getNormal(xlocsurf[0], xlocsurf[1])), surfobj.shape.lc)
    (kvector_4, efield_4) = sorted_unit_modes_isotropic(
        _isotropic_index(mat, wave), kconemat, surfnormalmat)
    xglob = lcobj.returnLocalToGlobalPoints(xlocobj)
    if not (np.any(kvector_4.imag) or np.any(efield_4.imag)):
        # real cones in a lossless medium: keep the bundle real-typed (the reference
        # carries complex128 with zero imaginary parts), so it runs in the real kernels
        (kvector_4, efield_4) = (kvector_4.real, efield_4.real)
    pilotbundles = []
    for j in range(4):
        kglob = mat.lc.returnLocalToGlobalDirections(kvector_4[j])
        eglob = mat.lc.returnLocalToGlobalDirections(efield_4[j])
        pilotbundles.append(RayBundle(x0=xglob, k0=kglob, Efield0=eglob, wave=wave))
    return pilotbundles


def build_pilotbundle(surfobj, mat, dxdy_pair, dphi_pair, efield_local_k=None,
                      kunitvector=None, lck=None, wave=standard_wavelength,
                      num_sampling_points=5, random_xy=False):
    """Real pilot bundle (reference :78-203): positions 0, +-dx.. x 0, +-dy.. on the
    object surface, directions on cones of half-angle 0 .. (dphix + dphiy)/2 about
    `kunitvector` (default: z of `lck`).  Returns the four sorted solutions as
    RayBundles; `[-1]` is a forward one."""
    (dx_val, dy_val) = dxdy_pair
    (phix, phiy) = dphi_pair
    lcobj = surfobj.rootcoordinatesystem
    if lck is None:
        lck = lcobj
    if kunitvector is None:
        kunitvector = np.array([0, 0, 1])
    lim_angle = 0.5 * (phix + phiy)
    n = num_sampling_points
    if not random_xy:
        lspace = _lspace(n)
        x_start = dx_val * lspace
        y_start = dy_val * lspace
    else:
        x_start = dx_val * np.hstack((0, 1. - 2. * np.random.random(n - 1)))
        y_start = dy_val * np.hstack((0, 1. - 2. * np.random.random(n - 1)))
    phi = np.arctan2(kunitvector[1], kunitvector[0])
    theta = np.arcsin(np.sqrt(kunitvector[1] ** 2 + kunitvector[0] ** 2))
    alpha = np.linspace(-lim_angle, 0, n, endpoint=False)
    angle = np.linspace(0, 2. * np.pi, n, endpoint=False)
    (alpha_grid, angle_grid, x_grid, y_grid) = np.meshgrid(alpha, angle, x_start, y_start)
    cone = np.vstack(((np.cos(angle_grid) * np.sin(alpha_grid)).flatten(),
                      (np.sin(angle_grid) * np.sin(alpha_grid)).flatten(),
                      np.cos(alpha_grid).flatten()))
    start_pts = np.vstack((x_grid.flatten(), y_grid.flatten(),
                           np.zeros_like(x_grid.flatten())))
    finalrot = np.dot(rodrigues(-theta, [1, 0, 0]), rodrigues(-phi, [0, 0, 1]))
    return _finish(surfobj, mat, start_pts, np.dot(finalrot, cone), lck, wave)


def build_pilotbundle_complex(surfobj, mat, dxdy_pair, dphi_pair, efield_local_k=None,
                              kunitvector=None, lck=None, wave=standard_wavelength,
                              num_sampling_points=3):
    """Complex pilot bundle (reference :206-316): a Cartesian raster in x, y, Re kx,
    Im kx, Re ky, Im ky, Im kz (num_sampling_points^7 rays) with Re kz from the unit
    Hermitian norm; the direction is always z of `lck` (the reference does not rotate
    it either, :262-276)."""
    (dx_val, dy_val) = dxdy_pair
    (phix, phiy) = dphi_pair
    lcobj = surfobj.rootcoordinatesystem
    if lck is None:
        lck = lcobj
    lim_angle = 0.5 * (phix + phiy)
    lspace = _lspace(num_sampling_points)
    kl = lim_angle * lspace
    (x_grid, y_grid, kxr, kxi, kyr, kyi, kzi) = np.meshgrid(
        dx_val * lspace, dy_val * lspace, kl, kl, kl, kl, kl)
    kzr = np.sqrt(1. - kxr ** 2 - kxi ** 2 - kyr ** 2 - kyi ** 2 - kzi ** 2)
    complex_ek = np.vstack((kxr.flatten() + 1j * kxi.flatten(),
                            kyr.flatten() + 1j * kyi.flatten(),
                            kzr.flatten() + 1j * kzi.flatten()))
    start_pts = np.vstack((x_grid.flatten(), y_grid.flatten(),
                           np.zeros_like(x_grid.flatten())))
    return _finish(surfobj, mat, start_pts, complex_ek, lck, wave)
