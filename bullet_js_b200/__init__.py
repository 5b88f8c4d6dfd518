"""bullet_js_b200 - B200-native drop-in for the bullet-js merge + index hot path.

(The task names the package `bullet-js_b200`; a hyphen cannot be imported, so the
directory is `bullet_js_b200`.)  Contents: `csrc/` (CUDA kernels + the C ABI of
include/bullet_b200.h), `capi` (ctypes loader), `codec` (host-side typed packing),
`engine` (thin object wrapper over the C ABI), `bullet` (host mirror of the
reference's Bullet / BulletNode / query interface for this path).
There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
