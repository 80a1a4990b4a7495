"""Paraxial callers of seqtrace: pilot bundles, XYUV transfer
matrices, linearised trace and Aimy against fixtures dumped from the unmodified
reference (oracle/gen_golden.py --paraxial).

CPU tests check the host logic: generators and matrix fits directly, and the whole
chain with the pilot traces served by the NumPy oracle (monkeypatched in place of
the native engine -- test infrastructure only).  GPU tests run the same chain
through the native engine."""
import os

import numpy as np
import pytest
import torch

import pyrate_b200 as pb
from pyrate_b200 import configs, engine
from pyrate_b200.raytracer import helpers, xyuv
from pyrate_b200.raytracer.aim import Aimy
from pyrate_b200.raytracer.optical_element import OpticalElement
from pyrate_b200.raytracer.ray import RayBundle, RayPath

import util

DEG = np.pi / 180.0
CASES = [("c2_doublegauss", 3.0), ("x1_tilted", 2.0), ("x7_two_elements", 3.0),
         ("c1_doublet", 4.0)]
GENERATORS = {"real": helpers.build_pilotbundle, "complex": helpers.build_pilotbundle_complex}
# matrices come out of normal equations with condition numbers ~1e4 (fixture log)
TOL_MATRIX = 1e-8


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(util.GOLDEN, "paraxial.npz"))


def _system(name):
    (s, seq) = configs.build_system(configs.CONFIGS[name], pb.api())
    objsurf = s.elements[seq[0][0]].surfaces[seq[0][1][0][0]]
    return (s, seq, objsurf)


def _pilot(s, objsurf, gen):
    return GENERATORS[gen](objsurf, s.material_background, (0.1, 0.1), (1 * DEG, 1 * DEG),
                           num_sampling_points=3)[-1]


def _close(a, b, tol, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))
    assert err <= tol, "%s: rel err %.3e > %.1e" % (what, err, tol)
    return err


@pytest.mark.parametrize("name", [c[0] for c in CASES])
@pytest.mark.parametrize("gen", ["real", "complex"])
def test_pilot_bundles_match_reference(gold, name, gen):
    (s, seq, objsurf) = _system(name)
    bundles = GENERATORS[gen](objsurf, s.material_background, (0.1, 0.1),
                              (1 * DEG, 1 * DEG), num_sampling_points=3)
    assert len(bundles) == 4
    d = bundles[-1].numpy()
    pre = "%s_%s_" % (name, gen)
    _close(d["x"][0], gold[pre + "pilot_x"], 1e-14, "pilot x")
    _close(d["k"][0], gold[pre + "pilot_k"], 1e-14, "pilot k")
    # E is arbitrary in the plane E.k = 0 (bilinear product), unit Hermitian norm
    (k, e) = (d["k"][0].astype(complex), d["Efield"][0].astype(complex))
    assert np.max(np.abs(np.sum(e * k, axis=0))) < 1e-14
    assert np.allclose(np.sum(np.abs(e) ** 2, axis=0), 1.0, atol=1e-14)
    # the backward pair is the mirror image of the forward pair
    assert np.allclose(bundles[0].numpy()["k"], -d["k"])
    if gen == "real":
        assert not bundles[-1].k.is_complex()


@pytest.mark.parametrize("name", ["c2_doublegauss", "x1_tilted"])
@pytest.mark.parametrize("gen", ["real", "complex"])
def test_pair_matrices_from_reference_pilot_path(gold, name, gen):
    """xyuv.transfer_matrices on the reference's own pilot hit points."""
    (s, seq, _) = _system(name)
    pre = "%s_%s_" % (name, gen)
    (elemkey, subseq) = seq[0]
    elem = s.elements[elemkey]
    (hitlist, _) = elem.sequence_to_hitlist(subseq)
    mats = xyuv.transfer_matrices(elem.surfaces, hitlist, gold[pre + "pilotpath_x"],
                                  gold[pre + "pilotpath_k"], gen)
    for (i, h) in enumerate(hitlist):
        _close(mats[h], gold[pre + "pair_matrices"][i], TOL_MATRIX, "pair %s" % (h,))
        _close(mats[(h[1], h[0], h[2])], gold[pre + "pair_inverse"][i], TOL_MATRIX,
               "inverse pair %s" % (h,))


def test_hitlist_round_trip():
    elem = OpticalElement.p(pb.LocalCoordinates.p(name="e"), name="e")
    seq = [("a", {"is_stop": True}), ("b", {}), ("a", {"is_mirror": True}), ("b", {}),
           ("c", {})]
    (hitlist, opts) = elem.sequence_to_hitlist(seq)
    assert hitlist == [("a", "b", 1), ("b", "a", 1), ("a", "b", 2), ("b", "c", 1)]
    assert opts[("a", "b", 2)] == ({"is_mirror": True}, {})
    back = elem.hitlist_to_sequence((hitlist, opts))
    assert [(k, o) for (k, _, o) in back] == seq


def test_choose_nearest():
    rng = np.random.default_rng(3)
    k4 = rng.normal(size=(4, 3, 6)) + 1j * rng.normal(size=(4, 3, 6))
    kvec = k4[1] + 1e-3 * rng.normal(size=(3, 6))
    kvec[:, 2] = k4[3][:, 2]                      # identical: excluded by the tolerance
    res = helpers.choose_nearest(kvec, k4)
    for j in range(6):
        d2 = [np.real(np.vdot(k4[i, :, j] - kvec[:, j], k4[i, :, j] - kvec[:, j]))
              for i in range(4)]
        cands = [i for i in range(4) if d2[i] > 1e-3]
        best = min(cands, key=lambda i: d2[i])
        assert np.array_equal(res[:, j], k4[best, :, j])


# ---------------------------------------------------------------------------
# whole chain; the pilot traces come from `tracer`
# ---------------------------------------------------------------------------
def _run_chain(gold, name, stopsize, gen):
    """extractXYUV, para_seqtrace and Aimy.aim on one system -> dict of arrays."""
    (s, seq, objsurf) = _system(name)
    pre = "%s_%s_" % (name, gen)
    out = {}
    (out["m_obj_stop"], out["m_stop_img"]) = s.extractXYUV(
        _pilot(s, objsurf, gen), seq, pilotbundle_generation=gen)
    x0 = gold[pre + "para_x0"]
    k0 = gold[pre + "para_k0"]
    (pp, rp) = s.para_seqtrace(_pilot(s, objsurf, gen),
                               RayBundle(x0, k0, None, wave=configs.DLINE), seq,
                               pilotbundle_generation=gen)
    out["para_x"] = np.array([b.numpy()["x"][-1] for b in rp.raybundles])
    out["para_k"] = np.array([b.numpy()["k"][-1] for b in rp.raybundles])
    a = Aimy(s, seq, wave=configs.DLINE, num_pupil_points=24, stopsize=stopsize,
             pilotbundle_generation=gen)
    out["aimy_m_obj_stop"] = a.m_obj_stop
    b1 = a.aim(np.array([0.01, -0.02]), fieldtype="angle").numpy()
    (out["aim_angle_x"], out["aim_angle_k"]) = (b1["x"][0], b1["k"][0])
    assert np.max(np.abs(np.sum(b1["Efield"][0] * b1["k"][0], axis=0))) < 1e-13
    try:
        b2 = a.aim(np.array([0.3, 0.1]), fieldtype="objectheight").numpy()
        (out["aim_height_x"], out["aim_height_k"]) = (b2["x"][0], b2["k"][0])
    except np.linalg.LinAlgError:       # stop = object surface: B block singular
        pass
    return out


def _check_real_chain(gold, name, stopsize, tol):
    """Real pilot bundles: everything against the unmodified reference."""
    got = _run_chain(gold, name, stopsize, "real")
    pre = "%s_real_" % name
    for key in ("m_obj_stop", "m_stop_img", "para_x", "para_k", "aimy_m_obj_stop",
                "aim_angle_x", "aim_angle_k", "aim_height_x", "aim_height_k"):
        assert (key in got) == (pre + key in gold), key
        if key in got:
            ref = gold[pre + key]
            _close(got[key], ref.real if not np.any(np.imag(ref)) else ref, tol, key)


def _check_complex_chain_sane(gold, got, name):
    """Complex pilot bundles cannot be pinned on the reference: downstream of the first
    refraction its E is an SVD null vector with conj(E).k = 0 instead of E.k = 0
    (material_isotropic.py:106-128), so the Poynting direction of a complex-k ray
    depends on LAPACK's arbitrary choice in a two-dimensional null space -- perturbing
    the pilot k by 1e-16 moves the reference's matrices by 4 %.
    The engine carries E with E.k = 0, which makes d = Re k / |Re k| and the matrices
    well defined: their x / Re k block must agree with the reference's REAL-pilot
    matrices up to the second-order terms of the fit, and must not couple to Im k."""
    for key in ("m_obj_stop", "m_stop_img"):
        m = got[key]
        ref = gold["%s_real_%s" % (name, key)]
        assert m.shape == (6, 6)
        _close(m[:4, :4], ref, 2e-3, key + " real block")
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(m[:4, 4:])) < 1e-9 * scale
        assert np.max(np.abs(m[4:, :4])) < 1e-9 * scale


def _oracle_element_seqtrace(spec):
    """Stand-in for OpticalElement.seqtrace served by the NumPy oracle."""
    import pyrate_np as onp
    system = onp.system_from_spec(spec)
    by_name = {st["name"]: st for st in system["steps"]}

    def seqtrace(self, raybundle, sequence, background_medium, splitup=False):
        steps = []
        for (surfkey, opts) in sequence:
            st = dict(by_name[surfkey])
            st["elem"] = 0
            st["is_mirror"] = bool(opts.get("is_mirror", False))
            steps.append(st)
        sub = {"background": system["background"], "steps": steps}
        d = raybundle.numpy()
        paths = onp.seqtrace(sub, d["x"][-1], d["k"][-1], d["Efield"][-1],
                             wave=raybundle.wave, splitup=splitup)
        out = []
        for p in paths:
            rp = RayPath()
            for b in p[1:]:
                rb = RayBundle(_lazy={"x": torch.from_numpy(b["x"]), "k": torch.from_numpy(b["k"]),
                                      "Efield": torch.from_numpy(b["E"]),
                                      "valid": torch.from_numpy(b["valid"]),
                                      "rayID": torch.from_numpy(b["rayID"])},
                               wave=raybundle.wave)
                rp.appendRayBundle(rb)
            out.append(rp)
        return out
    return seqtrace


def _oracle_engine(monkeypatch, name, bilinear_e=False):
    import pyrate_np as onp
    monkeypatch.setattr(OpticalElement, "seqtrace",
                        _oracle_element_seqtrace(configs.CONFIGS[name]))
    monkeypatch.setattr(engine, "compute_device", lambda: torch.device("cpu"))
    if bilinear_e:
        # the engine's E convention for complex k (E.k = 0, csrc/pyr_aniso.cu) in
        # place of the reference's SVD null vector
        monkeypatch.setattr(onp, "efield_svd", lambda k, eps: (
            helpers._perpendicular(k) if k.shape[1] else np.zeros_like(k)))


@pytest.mark.parametrize("name,stopsize", CASES)
def test_real_chain_with_oracle_traces(gold, monkeypatch, name, stopsize):
    _oracle_engine(monkeypatch, name)
    _check_real_chain(gold, name, stopsize, TOL_MATRIX)


@pytest.mark.parametrize("name,stopsize", CASES)
def test_complex_chain_with_oracle_traces(gold, monkeypatch, name, stopsize):
    _oracle_engine(monkeypatch, name, bilinear_e=True)
    _check_complex_chain_sane(gold, _run_chain(gold, name, stopsize, "complex"), name)


def test_para_seqtrace_needs_a_device(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    (s, seq, objsurf) = _system("c1_doublet")
    with pytest.raises(engine.DeviceRequired):
        s.extractXYUV(_pilot(s, objsurf, "real"), seq, pilotbundle_generation="real")


def test_lost_pilot_rays_raise(monkeypatch):
    """A pilot bundle that is clipped on its way cannot be fitted."""
    name = "c1_doublet"
    monkeypatch.setattr(OpticalElement, "seqtrace",
                        _oracle_element_seqtrace(configs.CONFIGS[name]))
    (s, seq, objsurf) = _system(name)
    wide = helpers.build_pilotbundle(objsurf, s.material_background, (30.0, 30.0),
                                     (1 * DEG, 1 * DEG), num_sampling_points=3)[-1]
    with pytest.raises(Exception, match="lost rays"):
        s.extractXYUV(wide, seq, pilotbundle_generation="real")


def test_two_stops_return_none():
    (s, seq, objsurf) = _system("c1_doublet")
    seq2 = [(seq[0][0], [(k, dict(o, is_stop=True)) for (k, o) in seq[0][1]])]
    assert s.extractXYUV(_pilot(s, objsurf, "real"), seq2) is None


# ---------------------------------------------------------------------------
# GPU: the same chain with the native engine tracing the pilot bundles
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,stopsize", CASES)
def test_real_chain_native(gold, name, stopsize):
    _check_real_chain(gold, name, stopsize, TOL_MATRIX)


@pytest.mark.gpu
@pytest.mark.parametrize("name,stopsize", CASES)
def test_complex_chain_native(gold, name, stopsize):
    """Complex pilot bundles (complex k through isotropic media: the complex-valued
    kernel) against the oracle with the engine's E convention, and the sanity bounds
    against the reference's real-pilot matrices."""
    got = _run_chain(gold, name, stopsize, "complex")
    _check_complex_chain_sane(gold, got, name)
    with pytest.MonkeyPatch.context() as mp:
        _oracle_engine(mp, name, bilinear_e=True)
        want = _run_chain(gold, name, stopsize, "complex")
    assert set(got) == set(want)
    for key in sorted(want):
        _close(got[key], want[key], 1e-7, key)


def test_element_analysis_calc_xyuv(gold, monkeypatch):
    name = "c2_doublegauss"
    _oracle_engine(monkeypatch, name)
    (s, seq, objsurf) = _system(name)
    (elemkey, subseq) = seq[0]
    elem = s.elements[elemkey]
    (hitlist, _) = elem.sequence_to_hitlist(subseq)
    got = pb.OpticalElementAnalysis(elem, subseq).calc_xyuv(
        hitlist[:6], _pilot(s, objsurf, "real"), subseq, s.material_background)
    _close(got, gold[name + "_real_m_obj_stop"], TOL_MATRIX, "obj -> stop product")
