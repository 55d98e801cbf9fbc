"""Convex-polyhedron membership used by the calib frustum filter (reference: det3d/core/bbox/geometry.py:215-275)."""
import numpy as np


def surface_equ_3d(polygon_surfaces):
    """Plane (normal, d) per surface from its first three points; normals point outwards for the reference's frusta."""
    v = polygon_surfaces[:, :, :2, :] - polygon_surfaces[:, :, 1:3, :]
    normal = np.cross(v[:, :, 0, :], v[:, :, 1, :])
    d = np.einsum("aij,aij->ai", normal, polygon_surfaces[:, :, 0, :])
    return normal, -d


def frustum_planes(polygon_surfaces):
    """[num_polygon, S, >=3, 3] -> [num_polygon, S, 4] (a, b, c, d) rows consumed by sessd_postprocess."""
    n, d = surface_equ_3d(np.asarray(polygon_surfaces)[:, :, :3, :])
    return np.concatenate([n, d[..., None]], axis=-1).astype(np.float32)


def points_in_convex_polygon_3d_jit(points, polygon_surfaces, num_surfaces=None):
    """[num_points, num_polygon] bool: inside iff a*x+b*y+c*z+d < 0 for every surface."""
    n, d = surface_equ_3d(np.asarray(polygon_surfaces)[:, :, :3, :])
    sign = np.einsum("pc,asc->pas", points, n) + d[None]
    return np.all(sign < 0, axis=-1)
