"""CPU: CatalogMaterial / IndexFormulaContainer against indices computed by the
reference from pages of the refractiveindex.info database (fixtures generated
by oracle/gen_golden.py --glasscat; the database itself is not shipped)."""
import json
import os

import numpy as np
import pytest

import pyrate_b200 as pb
from pyrate_b200.raytracer.material.material_glasscat import CatalogMaterial

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glasscat.json")
PAGES = json.load(open(FIX))


@pytest.mark.parametrize("page", PAGES, ids=[p["page"] for p in PAGES])
def test_catalogue_index_matches_reference(page):
    lc = pb.LocalCoordinates.p(name="gc")
    mat = CatalogMaterial.p(lc, {"DATA": page["DATA"]})
    for (w, nr, ni) in zip(page["waves_mm"], page["n_real"], page["n_imag"]):
        if abs(ni) > 1e-4 * abs(nr):
            with pytest.raises(NotImplementedError):
                mat.get_optical_index(None, w)
        else:
            assert np.isclose(mat.get_optical_index(None, w), nr, rtol=1e-14)
    with pytest.raises(Exception):
        mat.get_optical_index(None, 1e3)          # far outside the validity range


def test_formula_coverage():
    kinds = {p["DATA"][0]["type"] for p in PAGES}
    assert {"formula 1", "formula 2", "formula 3", "formula 4", "formula 5",
            "formula 6"} <= kinds


def test_catalogue_material_lowers_as_constant_index_medium():
    from pyrate_b200 import _native as nat, lowering
    lc = pb.LocalCoordinates.p(name="gc2")
    mat = CatalogMaterial.p(lc, {"DATA": PAGES[0]["DATA"]})
    m = lowering.lower_medium(mat, PAGES[0]["waves_mm"][1])
    assert m.kind == nat.MEDIUM_ISO_CONST
    assert np.isclose(m.n, PAGES[0]["n_real"][1], rtol=1e-14)


def test_retro_and_exotic_formulas():
    """Formulas 8 / 9 of the database (the reference declares them and raises
    NotImplementedError on evaluation, material_glasscat.py:403-413): the published
    expressions on the two database pages that use them, against literature indices
    (AgBr n_D = 2.253, Schroeter 1931; urea n_e(0.589 um) = 1.60, Rosker 1985) and against
    the expressions written out independently."""
    lc = pb.LocalCoordinates.p(name="gc3")
    agbr = CatalogMaterial.p(lc, {"DATA": [{"type": "formula 8", "wavelength_range": "0.495 0.67",
                                            "coefficients": "0.452505 0.09939 0.070537 -0.000150"}]})
    urea = CatalogMaterial.p(lc, {"DATA": [{"type": "formula 9", "wavelength_range": "0.3 1.06",
                                            "coefficients": "2.51527 0.0240 0.0300 0.020 1.52 0.8771"}]})
    w = 0.5893
    q = 0.452505 + 0.09939 * w ** 2 / (w ** 2 - 0.070537) - 0.000150 * w ** 2
    assert np.isclose(agbr.get_optical_index(None, w * 1e-3), np.sqrt((1 + 2 * q) / (1 - q)), rtol=1e-14)
    assert abs(agbr.get_optical_index(None, w * 1e-3) - 2.253) < 5e-3
    n2 = 2.51527 + 0.0240 / (w ** 2 - 0.0300) + 0.020 * (w - 1.52) / ((w - 1.52) ** 2 + 0.8771)
    assert np.isclose(urea.get_optical_index(None, w * 1e-3), np.sqrt(n2), rtol=1e-14)
    assert abs(urea.get_optical_index(None, w * 1e-3) - 1.60) < 1e-2
    with pytest.raises(Exception):
        agbr.get_optical_index(None, 0.4e-3)          # outside the page's validity range
