"""Synthetic Gaussian scenes and pinhole cameras of the BASELINE.json shapes.

No datasets are available offline, so every test, golden fixture and bench line is
driven from here (numpy only, explicit seeds, identical output on every machine).

Distributions follow SURVEY.md §8(d), which mirrors how SDP-GS itself initialises a
scene:
  * means       ~ U([-1,1]^3) * 1.3 * extent   (random-init cloud, scene/dataset_readers.py:553-555)
  * scales      ~ LogNormal(log(0.7 * (V_box/P)^(1/3)), 0.3)   (kNN-distance init, scene/gaussian_model.py:198-201)
  * rotations   = normalize(N(0,1)^4)            (activation at scene/gaussian_model.py:151)
  * opacity     = 0.1 ("init", scene/gaussian_model.py:205) or sigmoid(N(0,1.5)) ("trained")
  * SH          f_dc = RGB2SH(U(0,1)), f_rest ~ N(0,0.05), 16 coeffs/channel (utils/sh_utils.py:114-117)
  * feature     = normalize(N(0,1)^3)            (gaussian_renderer/__init__.py:283-287)
Cameras are built exactly the way scene/cameras.py:64-81 and utils/graphics_utils.py:38-84
build them: `world_view_transform` is the world-to-camera matrix TRANSPOSED,
`full_proj_transform = world_view_transform @ projection^T`, znear=0.01, zfar=100.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SH_C0 = 0.28209479177387814


@dataclass
class Camera:
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray  # f32[4,4], world->view, transposed (kernel reads element (r,c) at m[4c+r])
    projmatrix: np.ndarray  # f32[4,4], full projection, transposed
    campos: np.ndarray  # f32[3]


@dataclass
class Scene:
    means3D: np.ndarray  # f32[P,3]
    scales: np.ndarray  # f32[P,3]  (already activated: exp of the raw parameter)
    rotations: np.ndarray  # f32[P,4]  (unit quaternions, (r,x,y,z))
    opacities: np.ndarray  # f32[P,1]  (already activated: sigmoid)
    shs: np.ndarray  # f32[P,16,3]
    features: np.ndarray  # f32[P,3]
    confidence: np.ndarray  # f32[P,1]
    extent: float = 1.0
    cameras: list = field(default_factory=list)

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])


def look_at_camera(position, target, width, height, focal_px, znear=0.01, zfar=100.0) -> Camera:
    """COLMAP-convention camera (x right, y down, z forward) at `position` looking at `target`."""
    position = np.asarray(position, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    fwd = target - position
    fwd /= np.linalg.norm(fwd)
    up_hint = np.array([0.0, -1.0, 0.0])
    if abs(np.dot(up_hint, fwd)) > 0.99:
        up_hint = np.array([0.0, 0.0, 1.0])
    right = np.cross(up_hint, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    w2c = np.eye(4)
    w2c[0, :3], w2c[1, :3], w2c[2, :3] = right, down, fwd
    w2c[:3, 3] = -w2c[:3, :3] @ position
    w2c = np.float32(w2c)  # utils/graphics_utils.py:50 returns float32

    fovx = 2.0 * math.atan(width / (2.0 * focal_px))  # utils/graphics_utils.py:83-84
    fovy = 2.0 * math.atan(height / (2.0 * focal_px))
    tan_x, tan_y = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right_ = tan_y * znear, tan_x * znear
    proj = np.zeros((4, 4), dtype=np.float32)  # utils/graphics_utils.py:64-81
    proj[0, 0] = 2.0 * znear / (2.0 * right_)
    proj[1, 1] = 2.0 * znear / (2.0 * top)
    proj[3, 2] = 1.0
    proj[2, 2] = zfar / (zfar - znear)
    proj[2, 3] = -(zfar * znear) / (zfar - znear)

    view_t = np.ascontiguousarray(w2c.T)  # scene/cameras.py:78
    full = np.ascontiguousarray((view_t @ proj.T).astype(np.float32))  # scene/cameras.py:79-80
    campos = np.linalg.inv(view_t.astype(np.float64))[3, :3].astype(np.float32)  # scene/cameras.py:81
    return Camera(width, height, tan_x, tan_y, view_t, full, campos)


def ring_cameras(n_views, width, height, extent=1.0, radius=4.0, focal_frac=0.8, elevation=0.15, phase=0.0):
    cams = []
    for k in range(n_views):
        ang = phase + 2.0 * math.pi * k / max(n_views, 1)
        pos = np.array([radius * extent * math.cos(ang), elevation * radius * extent * math.sin(2 * ang + 0.3),
                        radius * extent * math.sin(ang)])
        cams.append(look_at_camera(pos, np.zeros(3), width, height, focal_frac * width))
    return cams


def make_scene(P, seed, extent=1.0, opacity="trained", scale_mult=1.0, sh_rest_std=0.05) -> Scene:
    rng = np.random.default_rng(seed)
    means = (rng.uniform(-1.0, 1.0, size=(P, 3)) * 1.3 * extent).astype(np.float32)
    v_box = (2.6 * extent) ** 3
    mu = math.log(0.7 * (v_box / max(P, 1)) ** (1.0 / 3.0) * scale_mult)
    scales = np.exp(rng.normal(mu, 0.3, size=(P, 3))).astype(np.float32)
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rotations = q.astype(np.float32)
    if opacity == "init":
        op = np.full((P, 1), 0.1, dtype=np.float32)
    else:
        op = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 1.5, size=(P, 1))))).astype(np.float32)
    shs = np.zeros((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = (rng.uniform(0.0, 1.0, size=(P, 3)) - 0.5) / SH_C0
    shs[:, 1:, :] = rng.normal(0.0, sh_rest_std, size=(P, 15, 3))
    f = rng.normal(size=(P, 3))
    f /= np.linalg.norm(f, axis=1, keepdims=True) + 1e-9
    return Scene(means, scales, rotations, op, shs.astype(np.float32), f.astype(np.float32),
                 np.ones((P, 1), dtype=np.float32), extent)


# BASELINE.md §4 configs as concrete synthetic inputs: (P, W, H, views)
CONFIGS = {
    "llff_fern_3view": dict(P=100_000, width=504, height=378, views=3, seed=1001),      # configs[0] / configs[1]
    "dtu_scan_3view": dict(P=300_000, width=400, height=300, views=3, seed=1003),       # configs[2]
    "mip360_render": dict(P=3_000_000, width=1297, height=840, views=200, seed=1004),   # configs[3]
    "stress_train": dict(P=6_000_000, width=1920, height=1080, views=8, seed=1005),     # configs[4]
    # small cases for parity tests / golden fixtures
    "tiny": dict(P=600, width=80, height=56, views=2, seed=7),
    "small": dict(P=4000, width=160, height=120, views=2, seed=11),
    # camera inside the cloud: near-plane culling, huge footprints, ragged image size (not a tile multiple)
    "inside": dict(P=3000, width=125, height=93, views=2, seed=13, radius=0.6),
}


def make_config(name, P=None, views=None, opacity="trained") -> Scene:
    c = dict(CONFIGS[name])
    if P is not None:
        c["P"] = P
    if views is not None:
        c["views"] = views
    sc = make_scene(c["P"], c["seed"], opacity=opacity)
    sc.cameras = ring_cameras(c["views"], c["width"], c["height"], extent=sc.extent, radius=c.get("radius", 4.0),
                              phase=0.1 * c["seed"])
    return sc


def cotangents(cam: Camera, seed):
    """dL/dcolor[3,H,W], dL/ddepth[1,H,W], dL/dalpha[1,H,W], dL/dfeature[3,H,W] ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    H, W = cam.height, cam.width
    return (rng.normal(size=(3, H, W)).astype(np.float32), rng.normal(size=(1, H, W)).astype(np.float32),
            rng.normal(size=(1, H, W)).astype(np.float32), rng.normal(size=(3, H, W)).astype(np.float32))
