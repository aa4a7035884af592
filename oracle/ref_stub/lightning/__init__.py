"""Import stub (test infrastructure): `lightning` is not installed in this image.

Only what the reference's hot-path modules touch at import/construct time is
provided, so that the UNMODIFIED files under /root/reference can be executed to
produce golden vectors (oracle/make_golden.py).  Never imported by the product.
"""
