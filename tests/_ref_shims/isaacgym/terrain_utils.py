"""Placeholder: the reference's Terrain class is monkeypatched by the harness."""
