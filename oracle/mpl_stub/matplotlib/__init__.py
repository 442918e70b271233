"""Test-infrastructure stub: lets the read-only reference import without matplotlib.

The reference's drawing/__init__.py:3-12 has an ImportError branch that does not define
Polygon / PatchCollection, so BeamFE2/Loads.py:4 fails to import when matplotlib is absent.
This stub provides just the names; any attempt to actually plot raises.
"""
