"""Catalogue glasses (host API mirror of reference
raytracer/material/material_glasscat.py: IndexFormulaContainer :290-470,
CatalogMaterial :473-538).

Dispersion is host-side scalar work -- one index per surface per wavelength --
so nothing here touches the device: `get_optical_index(x, wave)` feeds the
`n` of a PYR_MEDIUM_ISO_CONST medium at lowering time.  The refractiveindex.info
database itself (31 MB submodule of the reference) is not shipped; a material is
built from the parsed yml dict of one database page, exactly as
`CatalogMaterial.p(lc, ymldict)` of the reference.  Wavelengths are in mm on the
API, in um inside the formulas (the database's unit).
"""
import numpy as np

from .material_isotropic import IsotropicMaterial


def _pairs(c, start, step=2):
    return (c[start::step], c[start + 1::step])


def _n_formula1(c, w):          # Sellmeier
    (b, cc) = _pairs(c, 1)
    return np.sqrt(1 + c[0] + np.sum(b * w ** 2 / (w ** 2 - cc ** 2)))


def _n_formula2(c, w):          # Sellmeier, C already squared
    (b, cc) = _pairs(c, 1)
    return np.sqrt(1 + c[0] + np.sum(b * w ** 2 / (w ** 2 - cc)))


def _n_formula3(c, w):          # polynomial in n^2
    (a, p) = _pairs(c, 1)
    return np.sqrt(c[0] + np.sum(a * w ** p))


def _n_formula4(c, w):          # refractiveindex.info formula 4
    if len(c) > 10:
        (a, b, cc, d) = (c[[1, 5]], c[[2, 6]], c[[3, 7]], c[[4, 8]])
        (e, f) = (c[9::2], c[10::2])
        return np.sqrt(c[0] + np.sum(a * w ** b / (w ** 2 - cc ** d)) + np.sum(e * w ** f))
    (a, b, cc, d) = (c[1::4], c[2::4], c[3::4], c[4::4])
    return np.sqrt(c[0] + np.sum(a * w ** b / (w ** 2 - cc ** d)))


def _n_formula5(c, w):          # Cauchy
    (a, p) = _pairs(c, 1)
    return c[0] + np.sum(a * w ** p)


def _n_formula6(c, w):          # gases
    (b, cc) = _pairs(c, 1)
    return 1 + c[0] + np.sum(b / (cc - w ** (-2)))


def _n_formula7(c, w):          # Herzberger
    den = w ** 2 - 0.028
    a = c[3:]
    p = 2 * np.arange(len(a)) + 2
    return c[0] + c[1] / den + c[2] / den ** 2 + np.sum(a * w ** p)


def _n_formula8(c, w):
    """Retro: (n^2 - 1) / (n^2 + 2) = C1 + C2 w^2 / (w^2 - C3) + C4 w^2 (the database's
    "Dispersion formulas" document; the reference declares the type but raises
    NotImplementedError when it is evaluated, material_glasscat.py:403-407)."""
    c = np.concatenate((c, np.zeros(max(0, 4 - len(c)))))
    q = c[0] + c[1] * w ** 2 / (w ** 2 - c[2]) + c[3] * w ** 2
    return np.sqrt((1 + 2 * q) / (1 - q))


def _n_formula9(c, w):
    """Exotic: n^2 = C1 + C2 / (w^2 - C3) + C4 (w - C5) / ((w - C5)^2 + C6) (same document;
    NotImplementedError in the reference, :409-413)."""
    c = np.concatenate((c, np.zeros(max(0, 6 - len(c)))))
    return np.sqrt(c[0] + c[1] / (w ** 2 - c[2]) + c[3] * (w - c[4]) / ((w - c[4]) ** 2 + c[5]))


_FORMULAS = {"formula 8": _n_formula8, "formula 9": _n_formula9, "formula 1": _n_formula1, "formula 2": _n_formula2, "formula 3": _n_formula3,
             "formula 4": _n_formula4, "formula 5": _n_formula5, "formula 6": _n_formula6,
             "formula 7": _n_formula7}


class IndexFormulaContainer(object):
    """One dispersion entry of a database page (n or k)."""

    def __init__(self, typ, coeff, waverange):
        self.typ = typ
        self.coeff = np.asarray(coeff, dtype=float)
        self.waverange = np.asarray(waverange, dtype=float)
        if not (typ in _FORMULAS or typ.startswith("tabulated")):
            raise Exception("Bad dispersion function type: " + str(typ))

    def get_optical_index(self, wavelength):
        w = 1000.0 * wavelength                      # mm -> um
        if w < self.waverange[0] or w > self.waverange[1]:
            raise Exception("wavelength out of range: {0} um\nmust be between {1} um and "
                            "{2} um".format(w, *self.waverange))
        if self.typ in _FORMULAS:
            return _FORMULAS[self.typ](self.coeff, w)
        tab = self.coeff
        if self.typ == "tabulated n":
            return np.interp(w, tab[:, 0], tab[:, 1])
        if self.typ == "tabulated k":
            return 1j * np.interp(w, tab[:, 0], tab[:, 1])
        return np.interp(w, tab[:, 0], tab[:, 1]) + 1j * np.interp(w, tab[:, 0], tab[:, 2])


class CatalogMaterial(IsotropicMaterial):

    @classmethod
    def p(cls, lc, ymldict, name="", comment=""):
        return cls({"yml_dictionary": ymldict}, {"lc": lc, "comment": comment}, name=name)

    def setKind(self):
        self.kind = "material_from_catalog"

    def initialize_from_annotations(self):
        data = self.annotations["yml_dictionary"]["DATA"]
        if len(data) > 2:
            raise Exception("Max 2 entries for dispersion allowed - n and k.")
        self.nk_table = []
        for entry in data:
            typ = entry["type"]
            if typ.startswith("tabulated"):
                rows = [r.split() for r in entry["data"].split("\n") if r.strip()]
                coeff = np.array(rows, dtype=float)
                rang = np.array([coeff[:, 0].min(), coeff[:, 0].max()])
            else:
                coeff = np.array(entry["coefficients"].split(), dtype=float)
                rang = np.array(entry["wavelength_range"].split(), dtype=float)
            self.nk_table.append(IndexFormulaContainer(typ, coeff, rang))

    def get_optical_index(self, x, wave):
        n = 0
        for entry in self.nk_table:
            n = n + entry.get_optical_index(wave)
        # Database pages of optical glasses carry an extinction table next to the
        # dispersion formula (k ~ 1e-8): the reference then traces with a complex
        # index whose imaginary part only attenuates.  The real-valued device path
        # keeps Re(n) (the ray geometry changes by O(k^2)); media with a significant
        # extinction (metals) are refused instead of being traced as dielectrics.
        if np.iscomplexobj(n) and abs(np.imag(n)) > 1e-4 * abs(np.real(n)):
            raise NotImplementedError(
                "strongly absorbing catalogue material (n = %r): the isotropic device "
                "path is real-valued" % (n,))
        return float(np.real(n))
