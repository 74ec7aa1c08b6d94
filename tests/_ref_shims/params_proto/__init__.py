"""Minimal stand-in for the `params_proto` package (absent from this image).

Only the behaviour the reference relies on is provided (see neo_proto.py).
Test infrastructure: used solely by tests/golden/make_golden.py to import
/root/reference unmodified.
"""
