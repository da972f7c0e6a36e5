"""TEST INFRASTRUCTURE — shim package."""
