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

import math

import numpy as np

from .scene import Camera

FORWARD = np.array([0.0, 0.0, -1.0])
UP = np.array([0.0, 1.0, 0.0])


def rotation_from_direction(direction) -> np.ndarray:
    """mat3::from_3d_direction(normalize(dir)) (camera.hpp:116-119): a rotation whose forward axis is `dir`.  Plain float64
    arithmetic in the order the C++ host uses (scene_loader.hpp), so the two hosts produce the same matrix bit for bit."""
    def unit(v):
        length = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        return [v[0] / length, v[1] / length, v[2] / length]

    def cross(a, b):
        return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

    f = unit([float(x) for x in direction])
    back = [-f[0], -f[1], -f[2]]
    up = [0.0, 1.0, 0.0] if abs(f[1]) < 0.9999 else [0.0, 0.0, 1.0 if f[1] < 0 else -1.0]
    right = unit(cross(up, back))
    up2 = cross(back, right)
    return np.array([[right[r], up2[r], back[r]] for r in range(3)], np.float64)  # columns = images of +X, +Y, +Z


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
    """viewport::inverse_view_projection as 16 float32, column-major (element (r,c) at [c*4+r]).

    inverse(P * view) = world * inverse(P) with both factors written out -- world = translate(pos) * rot, and
    P^-1 = [[aspect/t,0,0,0],[0,1/t,0,0],[0,0,0,-1],[0,0,1/b,a/b]] -- instead of a numerical inverse: the structural zeros stay
    exact (the kernels take a cheaper perspective divide when the w row is (0, 0, m11, m15), which 1e-15 of inversion noise would
    switch off), and the C++ host (rt_b200/host/scene_loader.hpp) computes the same sums in the same order, so the two agree bit
    for bit."""
    rot = rotation_from_direction(cam.direction)
    pos = [float(x) for x in cam.position]
    world = [[float(rot[r, 0]), float(rot[r, 1]), float(rot[r, 2]), pos[r]] for r in range(3)] + [[0.0, 0.0, 0.0, 1.0]]
    t = 1.0 / math.tan(cam.vfov / 2.0)
    aspect = float(width) / float(height)
    n, fa = float(cam.near), float(cam.far)
    a, b = fa / (n - fa), n * fa / (n - fa)
    pinv = [[aspect / t, 0.0, 0.0, 0.0], [0.0, 1.0 / t, 0.0, 0.0], [0.0, 0.0, 0.0, -1.0], [0.0, 0.0, 1.0 / b, a / b]]
    out = np.zeros(16, np.float32)
    for r in range(4):
        for c in range(4):
            acc = 0.0
            for k in range(4):
                acc += world[r][k] * pinv[k][c]
            out[c * 4 + r] = np.float32(acc)
    return out
