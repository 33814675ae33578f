"""discover_tfbs_b200 — B200-native (sm_100a) feature-extraction hot path of discover_tfbs.

    sequence encode -> PWM log-odds scan (both strands) -> DNA-shape pentamer lookup -> features

Host code is Python over ctypes -> libtfbs_cuda.so (include/tfbs_cuda.h).  No CPU fallback.
Modules:  _lib (C-ABI binding)  utils (reference-signature mirror)  pssm (Biopython-shaped)
          dnashape (DNAshapeR-shaped)  sharding (multi-GPU chunk+halo)  synth (seeded inputs)
"""
__version__ = "0.1.0"
