"""CPU oracle for the spectral front-end hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product path
(``audio-deepfake-detection-fmsl_b200``) never imports it and fails loudly when its
CUDA library is missing.
"""
