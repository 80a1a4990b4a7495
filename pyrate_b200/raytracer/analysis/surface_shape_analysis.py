"""Sag tables of a shape (host API mirror of reference
raytracer/analysis/surface_shape_analysis.py:30-100; plotting left out).  Pure host
arithmetic on the shape's own getSag: a side tool of the path (e.g. to produce the
grid a `GridSag` surface is built from), not a part of it."""
import numpy as np


class ShapeAnalysis(object):

    def __init__(self, shape, name=""):
        self.shape = shape
        self.name = name

    def generate_sag_matrices(self, xlinspace, ylinspace):
        (xgrid, ygrid) = np.meshgrid(xlinspace, ylinspace)
        zflat = np.asarray(self.shape.getSag(xgrid.flatten(), ygrid.flatten()))
        return (xgrid, ygrid, np.reshape(zflat, np.shape(xgrid)))

    def generate_sag_table(self, xlinspace, ylinspace):
        (xgrid, ygrid, zgrid) = self.generate_sag_matrices(xlinspace, ylinspace)
        return np.vstack((xgrid.flatten(), ygrid.flatten(), zgrid.flatten()))

    def load_sag_table(self, filename):
        return np.loadtxt(filename, dtype=float).T

    def save_sag_table(self, filename, xlinspace, ylinspace):
        np.savetxt(filename, self.generate_sag_table(xlinspace, ylinspace).T)

    def compare_with_sag_table(self, table):
        return np.asarray(self.shape.getSag(table[0], table[1])) - table[2]
