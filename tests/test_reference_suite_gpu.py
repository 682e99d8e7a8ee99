"""The reference's own test file (reference tests/test_optimize.py:10-99), call for call and assertion for assertion,
run against the drop-in package: same positional arguments, same checks on the returned objects.  (The numerical
parity of the same four runs is what tests/test_gpu_parity.py checks against the goldens.)"""
import unittest

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.usefixtures("lib_built")
class TestSolidsPyOpt(unittest.TestCase):
    def setUp(self):
        import solidspy_opt.optimize                      # the drop-in package (tests/conftest.py puts it on sys.path)
        self.opt = solidspy_opt.optimize
        # Common parameters for the tests (reference tests/test_optimize.py:13-26)
        self.length = 60
        self.height = 60
        self.nx = 60
        self.ny = 60
        self.dirs = np.array([[0, -1]])
        self.positions = np.array([[15, 1]])
        self.niter = 50
        self.penal = 3
        self.volfrac = 0.5
        self.RR = 0.005
        self.ER = 0.05
        self.t = 1e-3

    def test_SIMP(self):
        rho = self.opt.SIMP(self.length, self.height, self.nx, self.ny, self.dirs, self.positions, self.niter,
                            self.penal, plot=False)
        self.assertIsInstance(rho, np.ndarray, "SIMP should return a numpy array.")
        self.assertEqual(rho.size, self.nx*self.ny, "Density array size should match the number of elements.")

    def test_ESO_stress(self):
        els, nodes = self.opt.ESO_stress(self.length, self.height, self.nx, self.ny, self.dirs, self.positions,
                                         self.niter, self.RR, self.ER, self.volfrac, plot=False)
        self.assertIsInstance(els, np.ndarray, "ESO_stress should return an array of elements.")
        self.assertIsInstance(nodes, np.ndarray, "ESO_stress should return an array of nodes.")
        self.assertTrue(els.size > 0, "The element array should not be empty.")

    def test_ESO_stiff(self):
        els, nodes = self.opt.ESO_stiff(self.length, self.height, self.nx, self.ny, self.dirs, self.positions,
                                        self.niter, self.RR, self.ER, self.volfrac, plot=False)
        self.assertIsInstance(els, np.ndarray, "ESO_stiff should return an array of elements.")
        self.assertIsInstance(nodes, np.ndarray, "ESO_stiff should return an array of nodes.")
        self.assertTrue(els.size > 0, "The element array should not be empty.")

    def test_BESO(self):
        els, nodes = self.opt.BESO(self.length, self.height, self.nx, self.ny, self.dirs, self.positions, self.niter,
                                   self.t, self.ER, self.volfrac, plot=False)
        self.assertIsInstance(els, np.ndarray, "BESO should return an array of elements.")
        self.assertIsInstance(nodes, np.ndarray, "BESO should return an array of nodes.")
        self.assertTrue(els.size > 0, "The element array should not be empty.")
