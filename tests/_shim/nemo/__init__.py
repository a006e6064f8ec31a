"""Test-only stand-in for the `nemo` package (absent from this image).

Only the handful of names the reference's `modules/`, `parts/` and `loss/` import at module
load time are provided; none of them affects arithmetic.  Never imported by the product.
"""
