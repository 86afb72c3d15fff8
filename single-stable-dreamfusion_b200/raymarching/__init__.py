"""Drop-in for the reference's ``raymarching`` package (raymarching/raymarching.py)."""
from .raymarching import (near_far_from_aabb, sph_from_ray, morton3D, morton3D_invert, packbits,  # noqa: F401
                          march_rays_train, composite_rays_train, march_rays, composite_rays, compact_alive)
