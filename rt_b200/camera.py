"""Viewport matrices: the host-side restatement of rt::camera::viewport (reference src/camera.hpp:122-137).

In a real integration the plugin passes muu's own `viewport::inverse_view_projection` through the C
ABI (plugin/cuda_path_tracer.cpp); the harness has no muu, so this module rebuilds the same matrix
from the camera pose.  Conventions (UNVERIFIED against muu@06dbcecb): right-handed,
forward = -Z, up = +Y, column-major storage, clip-space depth in [0,1] (depth 0 = near plane, 1 = far
plane, as `screen_to_world(pos, 0.0f / 1.0f)` at mg_ray_tracer.cpp:190-191 implies).

The matrix is *input data* to both the oracle and the CUDA path (they receive the same 16 floats), so
it is computed in float64 and rounded once to float32.
"""
from __future__ import annotations

import numpy as np

from .scene import Camera

FORWARD = np.array([0.0, 0.0, -1.0])
UP = np.array([0.0, 1.0, 0.0])


def rotation_from_direction(direction) -> np.ndarray:
    """mat3::from_3d_direction(normalize(dir)) (camera.hpp:116-119): a rotation whose forward axis is `dir`."""
    f = np.asarray(direction, np.float64)
    f = f / np.linalg.norm(f)
    back = -f
    up = UP if abs(np.dot(f, UP)) < 0.9999 else np.array([0.0, 0.0, 1.0 if f[1] < 0 else -1.0])
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    up2 = np.cross(back, right)
    return np.stack([right, up2, back], axis=1)  # columns = images of +X, +Y, +Z


def perspective_projection(vfov: float, width: int, height: int, near: float, far: float) -> np.ndarray:
    """mat4::perspective_projection(vfov, vec2{size}, near, far): RH, depth 0..1 (camera.hpp:131)."""
    f = 1.0 / np.tan(vfov / 2.0)
    aspect = float(width) / float(height)
    p = np.zeros((4, 4))
    p[0, 0] = f / aspect
    p[1, 1] = f
    p[2, 2] = far / (near - far)
    p[2, 3] = near * far / (near - far)
    p[3, 2] = -1.0
    return p


def view_matrix(cam: Camera) -> np.ndarray:
    """invert(from_translation(pos) * from_3d_rotation(rot)) (camera.hpp:130)."""
    world = np.eye(4)
    world[:3, :3] = rotation_from_direction(cam.direction)
    world[:3, 3] = np.asarray(cam.position, np.float64)
    return np.linalg.inv(world)


def inverse_view_projection(cam: Camera, width: int, height: int) -> np.ndarray:
    """viewport::inverse_view_projection as 16 float32, column-major (element (r,c) at [c*4+r])."""
    vp = perspective_projection(cam.vfov, width, height, cam.near, cam.far) @ view_matrix(cam)
    inv = np.linalg.inv(vp)
    return np.ascontiguousarray(inv.T.astype(np.float32).reshape(16))
