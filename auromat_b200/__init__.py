"""auromat_b200 -- B200-native (sm_100a) georeferencing + regridding path of esa/auromat.

    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample
    m = getMapping(img, header, altitude=110)        # lazy; coordinate planes live in HBM
    r = resample(m, arcsecPerPx=100, method='mean')  # CUDA binning + normalisation

The module layout follows the reference (`auromat.mapping.spacecraft`, `auromat.resample`,
`auromat.coordinates.*`).  The host side is Python and talks to hand-written CUDA kernels
through the C ABI in `include/auromat_b200.h` (ctypes).  There is no CPU fallback.
"""
__version__ = '0.1.0'
