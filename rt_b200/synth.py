"""Synthetic benchmark scenes of BASELINE.json (configs 3-5), SURVEY.md section 8d.

The generators are deterministic (numpy default_rng with the seeds fixed in SURVEY.md) and return a
`Scene`; `rt_b200.scene.dumps` turns one into a TOML file the reference loader accepts, which is how
`rt --scene <toml> --renderer cuda_path_tracer` would receive it.
"""
from __future__ import annotations

import numpy as np

from .scene import Camera, DIELECTRIC, LAMBERT, METAL, Scene, make_materials


def rtiow_scene(seed: int = 20260118, samples_per_pixel: int = 256, max_bounces: int = 50) -> Scene:
    """C3 / C5: 'Ray Tracing in One Weekend' final scene, ~488 spheres, mixed materials.

    Ground sphere (0,-1000,0) r=1000 lambert 0.5 gray; for a,b in [-11,11): centre (a+0.9u, 0.2, b+0.9u)
    r=0.2, skipped within 0.9 of (4,0.2,0); material by u: <0.8 lambert albedo=u*u, <0.95 metal albedo
    U[.5,1] roughness U[0,.5], else glass IOR 1.5; three r=1 spheres (glass, lambert, metal); camera
    (13,2,3) looking at the origin.  Glass albedo is 1/1.5 so that the reference's attenuation =
    albedo * reflectivity(=IOR) quirk (sm_ray_tracer.cpp:194) yields a neutral 1.0.
    """
    rng = np.random.default_rng(seed)
    mats = [(LAMBERT, (0.5, 0.5, 0.5), 0.5, 1.0)]
    spheres = [(0.0, -1000.0, 0.0, 1000.0)]
    sphere_mat = [0]
    glass = (DIELECTRIC, (1 / 1.5, 1 / 1.5, 1 / 1.5), 0.0, 1.5)
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = rng.random()
            c = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            u = rng.random(6)
            if np.linalg.norm(np.array(c) - np.array((4.0, 0.2, 0.0))) <= 0.9:
                continue
            if choose < 0.8:
                m = (LAMBERT, tuple(u[0:3] * u[3:6]), 0.5, 1.0)
            elif choose < 0.95:
                m = (METAL, tuple(0.5 + 0.5 * u[0:3]), 0.5 * u[3], 1.0)
            else:
                m = glass
            mats.append(m)
            spheres.append((*c, 0.2))
            sphere_mat.append(len(mats) - 1)
    for centre, m in (((0.0, 1.0, 0.0), glass), ((-4.0, 1.0, 0.0), (LAMBERT, (0.4, 0.2, 0.1), 0.5, 1.0)),
                      ((4.0, 1.0, 0.0), (METAL, (0.7, 0.6, 0.5), 0.0, 1.0))):
        mats.append(m)
        spheres.append((*centre, 1.0))
        sphere_mat.append(len(mats) - 1)
    return Scene(samples_per_pixel=samples_per_pixel, max_bounces=max_bounces,
                 camera=Camera(position=(13.0, 2.0, 3.0), direction=(-13.0, -2.0, -3.0)),
                 materials=make_materials(mats), material_names=[""] * len(mats),
                 spheres=np.array(spheres, np.float32), sphere_material=np.array(sphere_mat, np.uint32))


def grid_scene(nx: int = 400, nz: int = 250, seed: int = 20260119, samples_per_pixel: int = 64, max_bounces: int = 10) -> Scene:
    """C4: ground + nx*nz spheres on a jittered grid (spacing 0.5, r in U[0.08,0.22], y = r,
    x in [-nx/4, nx/4), z in [-nz/2, 0)), materials 70/20/10 lambert/metal/dielectric from a 64-entry
    palette; camera (0,6,8) direction (0,-0.35,-1).  Defaults give 100 001 spheres."""
    rng = np.random.default_rng(seed)
    mats = [(LAMBERT, (0.5, 0.5, 0.5), 0.5, 1.0)]
    for _ in range(44):
        mats.append((LAMBERT, tuple(0.1 + 0.8 * rng.random(3)), 0.5, 1.0))
    for _ in range(13):
        mats.append((METAL, tuple(0.5 + 0.5 * rng.random(3)), 0.3 * rng.random(), 1.0))
    for _ in range(6):
        mats.append((DIELECTRIC, (1 / 1.5, 1 / 1.5, 1 / 1.5), 0.0, 1.5))
    n = nx * nz
    ix, iz = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    r = rng.uniform(0.08, 0.22, n)
    x = -nx / 4 + 0.5 * (ix.ravel() + 0.15 + 0.2 * rng.random(n))
    z = -nz / 2 + 0.5 * (iz.ravel() + 0.15 + 0.2 * rng.random(n))
    kind = rng.random(n)
    mat = np.where(kind < 0.7, 1 + rng.integers(0, 44, n), np.where(kind < 0.9, 45 + rng.integers(0, 13, n), 58 + rng.integers(0, 6, n)))
    spheres = np.concatenate([np.array([[0.0, -1000.0, 0.0, 1000.0]]), np.stack([x, r, z, r], axis=1)]).astype(np.float32)
    sphere_mat = np.concatenate([[0], mat]).astype(np.uint32)
    return Scene(samples_per_pixel=samples_per_pixel, max_bounces=max_bounces,
                 camera=Camera(position=(0.0, 6.0, 8.0), direction=(0.0, -0.35, -1.0)),
                 materials=make_materials(mats), material_names=[""] * len(mats), spheres=spheres, sphere_material=sphere_mat)


def random_rays(scene: Scene, n: int, seed: int = 1, spread: float = 12.0) -> tuple[np.ndarray, np.ndarray]:
    """Seeded ray batch for level-1 parity: origins scattered around (and some on/inside) the scene's
    spheres, unit directions; a quarter of the rays start on a sphere surface pointing inward or outward."""
    rng = np.random.default_rng(seed)
    o = rng.normal(0.0, 1.0, (n, 3)) * np.array([spread, spread / 4, spread]) + np.array([0.0, 2.0, 0.0])
    d = rng.normal(0.0, 1.0, (n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if len(scene.spheres):
        k = n // 4
        idx = rng.integers(0, len(scene.spheres), k)
        sp = scene.spheres[idx].astype(np.float64)
        nrm = rng.normal(0.0, 1.0, (k, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        scale = np.where(rng.random(k) < 0.3, 0.5, 1.0)[:, None]  # some origins strictly inside
        o[:k] = sp[:, :3] + nrm * sp[:, 3:4] * scale
    return o.astype(np.float32), d.astype(np.float32)


def grazing_rays(scene: Scene, n: int, seed: int = 3) -> tuple[np.ndarray, np.ndarray]:
    """Adversarial batch for traversal parity: rays aimed at points on (or a hair inside / outside) the silhouette of
    randomly chosen spheres, from origins 0.3 to 400 units away, plus origins exactly on sphere surfaces.  These are the
    rays for which the rounding of the sphere test decides hit or miss."""
    rng = np.random.default_rng(seed)
    sp = scene.spheres[rng.integers(0, len(scene.spheres), n)].astype(np.float64)
    c, r = sp[:, :3], sp[:, 3:4]
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)      # direction origin -> centre
    dist = r + np.exp(rng.uniform(np.log(0.3), np.log(400.0), (n, 1)))
    o = c - u * dist
    # a unit vector perpendicular to u
    w = np.cross(u, rng.normal(size=(n, 3))); w /= np.linalg.norm(w, axis=1, keepdims=True)
    offs = rng.choice([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3], (n, 1))
    # tangent point direction: angle between u and the tangent line is asin(r / dist)
    sin_a = np.clip(r * (1.0 + offs) / dist, -1.0, 1.0)
    cos_a = np.sqrt(1.0 - sin_a ** 2)
    d = u * cos_a + w * sin_a
    k = n // 8  # some origins on the surface, pointing in or out
    o[:k] = c[:k] + w[:k] * r[:k]
    d[:k] = rng.normal(size=(k, 3)); d[:k] /= np.linalg.norm(d[:k], axis=1, keepdims=True)
    return o.astype(np.float32), (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
