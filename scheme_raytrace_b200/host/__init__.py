"""Host-side mirror of the reference's scene-construction API (one module per reference file)."""
