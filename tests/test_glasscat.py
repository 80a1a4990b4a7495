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
