"""vgpu-manager_b200 - B200-native drop-in for coldzerofear/vgpu-manager's libvgpu-control.so.

The product is the C library `vgpu_manager_b200/libvgpu-control.so` (LD_PRELOAD surface +
`vgpu_b200_*` direct entry points, see include/vgpu_b200.h).  This Python package only builds
it (`build`) and binds the direct entry points with ctypes (`lib`) for tests and bench.py.
"""
from .build import build, OUT as LIBRARY_PATH  # noqa: F401
from .lib import B200Library, LibraryMissing  # noqa: F401
