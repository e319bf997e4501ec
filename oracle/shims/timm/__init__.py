"""Empty stand-in for `timm` (TEST INFRASTRUCTURE, not product code).

The reference imports `timm` at model/trajectory_model.py:6 but never uses it; the
package is absent from this image and cannot be installed offline.
"""
