"""Test-infrastructure stand-in for the third-party ``solidspy`` package (>= 1.1 API).

ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

The reference (SolidsPy-Opt) imports ``solidspy.assemutil``, ``solidspy.postprocesor``,
``solidspy.uelutil`` and ``solidspy.preprocesor`` (reference ``src/solidspy_opt/optimize.py:10-11``,
``Utils/solver.py:4``, ``Utils/beams.py:2``) but the dependency is un-pinned (``pyproject.toml:31``),
not vendored, and not installable here (no network).  This package restates the *published
behaviour* of the seven functions the reference calls (contract: SURVEY.md Appendix A) so that
the untouched reference source can be executed to generate golden vectors.

Pinned against the known answers the reference tree does hold:
  * ``docs/tutorials/square_example.rst:150,187`` (2x2 square, 14 equations, displacement vector)
  * ``examples/beam_convergence/error_vs_h.txt`` (indirectly, through rect_grid meshes)
see ``tests/test_oracle_compat.py``.
"""
__version__ = "1.1.0+oracle-compat"
