"""Host-side, per-frame scalar work of the georeferencing path (WGS84 constants, IGRF dipole,
frame rotation matrices, WCS header digestion).  Everything here runs once per frame in
microseconds; the per-pixel passes are CUDA kernels (see auromat_b200/csrc)."""
