"""Empty stand-in for matplotlib (import-only use at curriculum.py:2)."""
