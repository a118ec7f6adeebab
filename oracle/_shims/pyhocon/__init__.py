"""
TEST INFRASTRUCTURE ONLY.  Minimal ``pyhocon`` stand-in so that the read-only reference at
/root/reference can be imported in the build container (pyhocon is not installed and there
is no network).  It simply re-exports this repo's HOCON-subset reader, which implements the
accessor surface the reference uses (src/util/args.py:6,90-101; nerf.py:340-352).
"""
from pixel_nerf_multiscale_b200.util.conf import ConfigFactory, ConfigTree  # noqa: F401
