"""TEST INFRASTRUCTURE — shim package (see utilities/__init__.py)."""
