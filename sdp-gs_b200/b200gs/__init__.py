"""b200gs: Blackwell-native (sm_100a) differentiable Gaussian-splatting rasterizer -- host package.

Layout: `_lib` (ctypes binding of the C-ABI in include/b200gs.h), `rasterizer` (autograd op),
`synthetic` (scenes/cameras of the BASELINE shapes), `bytes_model` (algorithmic byte model of
SURVEY.md Appendix E), `parallel` (view-sharded rendering / image-parallel training).
Importing `b200gs.rasterizer` requires the built shared library; `synthetic` and `bytes_model`
are numpy-only.
"""
