"""Linear (XYUV) transfer matrices from a traced pilot bundle, and their application
to a bundle of parabasal rays -- the arithmetic of reference
raytracer/optical_element.py:165-322 (`calculateXYUV`) and :381-469 (`para_seqtrace`).

The pilot bundle is traced by the native engine like any other bundle; what is left
here is a least-squares fit of a 4x4 / 6x6 matrix per surface pair on <= a few
thousand pilot rays (host NumPy on copies of the tiny pilot records) and, for
`para_seqtrace`, one small matrix product per surface over the parabasal bundle
(torch, on whatever device the bundle lives).
"""
import numpy as np


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def reduce_matrix(m):
    """Subtract the pilot ray (column 0) and keep the x, y rows of the other rays
    (reference :173-179)."""
    m = _np(m)
    return np.array((m - m[:, 0].reshape((3, 1)))[0:2, 1:])


def generate_matrix(x, k, generation="complex"):
    """(6, N-1) = [dx, dy, Re dkx, Re dky, Im dkx, Im dky] ("complex") or the first
    four rows ("real") -- reference generate_matrix_6xN :213-221."""
    xred = reduce_matrix(x).real
    kred = reduce_matrix(k)
    if generation.lower() == "complex":
        return np.vstack((xred, kred.real, kred.imag))
    return np.vstack((xred, kred.real))


def bestfit_transfer(xmat, ymat):
    """Least-squares T with Y ~ T X through the normal equations, exactly as the
    reference does (:201-211): T = (Y X^T) (X X^T)^-1."""
    xx = np.einsum('ij, kj', xmat, xmat).T
    yx = np.einsum('ij, kj', xmat, ymat).T
    return np.dot(yx, np.linalg.inv(xx))


def sequence_to_hitlist(seq):
    """Reference optical_element.py:128-151: consecutive surface pairs with a hit
    counter, plus their option dicts."""
    surfnames = [(name, options_dict) for (name, options_dict) in seq]
    hitlist_dict = {}
    hitlist = []
    optionshitlistdict = {}
    for ((sb, optsb), (se, optse)) in zip(surfnames[:-1], surfnames[1:]):
        hit = hitlist_dict.get((sb, se), 0) + 1
        hitlist_dict[(sb, se)] = hit
        hitlist.append((sb, se, hit))
        optionshitlistdict[(sb, se, hit)] = (optsb, optse)
    return (hitlist, optionshitlistdict)


def hitlist_to_sequence(hitlist_pair):
    """Reference optical_element.py:153-163."""
    (hitlist, optionshitlistdict) = hitlist_pair
    seq = []
    for (ind, (sb, se, hit)) in enumerate(hitlist):
        seq.append((sb, True, optionshitlistdict[(sb, se, hit)][0]))
        if ind == len(hitlist) - 1:
            seq.append((se, True, optionshitlistdict[(sb, se, hit)][1]))
    return seq


def transfer_matrices(surfaces, hitlist, pilot_x, pilot_k, generation="complex"):
    """XYUV matrices of one element.

    surfaces: dict key -> Surface; hitlist: [(s1, s2, hit), ...];
    pilot_x[i], pilot_k[i]: (3, N) GLOBAL hit point / wave vector of the pilot bundle
    arriving at the i-th surface of the sequence (the last rows of the path's bundles,
    reference :268-289).  Returns {(s1, s2, hit): T, (s2, s1, hit): T_inverse_fit}."""
    out = {}
    for (i, (s1, s2, numhit)) in enumerate(hitlist):
        lcstart = surfaces[s1].rootcoordinatesystem
        lcend = surfaces[s2].rootcoordinatesystem
        startx = lcstart.returnGlobalToLocalPoints(_np(pilot_x[i]))
        startk = lcstart.returnGlobalToLocalDirections(_np(pilot_k[i]))
        endx = lcend.returnGlobalToLocalPoints(_np(pilot_x[i + 1]))
        endk = lcend.returnGlobalToLocalDirections(_np(pilot_k[i + 1]))
        startmatrix = generate_matrix(startx, startk, generation)
        endmatrix = generate_matrix(endx, endk, generation)
        out[(s1, s2, numhit)] = bestfit_transfer(startmatrix, endmatrix)
        out[(s2, s1, numhit)] = bestfit_transfer(endmatrix, startmatrix)
    return out


def para_step(lc_start, lc_end, matrix, x0_glob, k0_glob, px0_glob, pk0_glob,
              px1_glob, pk1_glob, generation="complex"):
    """One surface pair of `para_seqtrace` (reference :401-457) in torch:
    (x0, k0) global (3, N) of the parabasal bundle at the start surface, pilot ray
    (3,) at both surfaces; returns global (x1, k1) at the end surface."""
    import torch
    dev = x0_glob.device
    cplx = generation.lower() == "complex"

    def t(a, dtype):
        a = np.asarray(a)
        if np.iscomplexobj(a) and dtype == torch.float64:
            a = a.real
        return torch.as_tensor(a, dtype=dtype, device=dev)

    kdt = torch.complex128 if (cplx or k0_glob.is_complex()) else torch.float64
    x0 = lc_start.returnGlobalToLocalPoints(x0_glob)
    k0 = lc_start.returnGlobalToLocalDirections(k0_glob.to(kdt))
    px0 = lc_start.returnGlobalToLocalPoints(t(px0_glob, torch.float64).reshape(3, 1))
    pk0 = lc_start.returnGlobalToLocalDirections(t(pk0_glob, kdt).reshape(3, 1))
    px1 = lc_end.returnGlobalToLocalPoints(t(px1_glob, torch.float64).reshape(3, 1))
    pk1 = lc_end.returnGlobalToLocalDirections(t(pk1_glob, kdt).reshape(3, 1))
    dx0 = (x0 - px0)[0:2]
    dk0 = (k0 - pk0)[0:2]
    if cplx:
        big = torch.cat((dx0, dk0.real, dk0.imag))
    else:
        big = torch.cat((dx0, dk0.real if dk0.is_complex() else dk0))
    res = t(matrix, torch.float64) @ big
    n = res.shape[1]
    zeros = torch.zeros((1, n), dtype=torch.float64, device=dev)
    dx1 = torch.cat((res[0:2], zeros))
    if cplx:
        dk1 = torch.cat((torch.complex(res[2:4], res[4:6]), zeros.to(torch.complex128)))
    else:
        dk1 = torch.cat((res[2:4], zeros)).to(kdt)
    x1 = lc_end.returnLocalToGlobalPoints(dx1 + px1)
    k1 = lc_end.returnLocalToGlobalDirections(dk1 + pk1)
    return (x1, k1)
