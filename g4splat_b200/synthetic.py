"""Seeded synthetic surfel scenes and cameras for tests and benchmarks.

"Room" scene: surfels on the inner faces of a 6 x 3 x 6 m box centred at the origin plus 10 % on
three interior spheres; cameras walk the perimeter of the room and look across it.  Stands in
for the indoor scans named by BASELINE.json (Replica / ScanNet++ / DeepBlending), none of which
is available offline.  Two statistics were tuned against the CPU oracle so that the workload
looks like a converged scan rather than a sparse point cloud (SURVEY.md 8d guessed both):
  * surfel sigma = 1.0 * sqrt(area / P) (the mean surfel spacing): accumulated alpha ~ 0.99 on
    surfaces and ~130 list entries visited per pixel; with the 0.5 factor first proposed the
    walls are 30 % transparent and the blend stages are almost idle;
  * perimeter cameras see ~25 % of the surfels (a camera in the room centre with a 60 degree
    lens sees 3-13 %).
Everything is generated on the CPU in float32 from `numpy.random.default_rng(seed)` so that the
oracle, the reference extension and the B200 kernels see bit-identical inputs.

Camera matrices follow the conventions of the reference's scene/cameras.py:55-58 and
utils/graphics_utils.py:38-71: `viewmatrix` is the TRANSPOSED world-to-camera matrix,
`projmatrix` the transposed (projection @ world-to-camera), camera looks down +z, y down.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

BOX = np.array([6.0, 3.0, 6.0], dtype=np.float64)  # x, y (height), z extents
SPHERES = np.array([[1.2, -0.9, 0.8], [-1.5, -0.7, -1.0], [0.3, 0.4, -1.8]], dtype=np.float64)
SPHERE_R = 0.5
SH_C0 = 0.28209479177387814

# (P, W, H, number of cameras, seed) of the BASELINE.json configs c0..c4
CONFIGS = {
    "c0": dict(P=10_000, W=256, H=256, cams=1, seed=0),
    "c1": dict(P=200_000, W=1200, H=680, cams=5, seed=1),
    "c2": dict(P=1_000_000, W=1920, H=1080, cams=1, seed=2),
    "c3": dict(P=2_500_000, W=1600, H=1200, cams=50, seed=3),
    "c4": dict(P=5_000_000, W=1920, H=1080, cams=64, seed=4),
}


def _normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _frame_quaternion(normal: np.ndarray, angle: np.ndarray) -> np.ndarray:
    """(w,x,y,z) quaternions of rotations whose third column is `normal`, spun by `angle`."""
    n = _normalize(normal)
    helper = np.where(np.abs(n[:, :1]) < 0.9, np.array([[1.0, 0.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))
    u = _normalize(np.cross(helper, n))
    v = np.cross(n, u)
    c, s = np.cos(angle)[:, None], np.sin(angle)[:, None]
    u2, v2 = c * u + s * v, -s * u + c * v
    R = np.stack([u2, v2, n], axis=2)  # columns
    # rotation matrix -> quaternion (numerically safe branch selection)
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    q = np.empty((R.shape[0], 4))
    tr = m00 + m11 + m22
    w = np.sqrt(np.maximum(0.0, 1.0 + tr)) / 2
    x = np.sqrt(np.maximum(0.0, 1.0 + m00 - m11 - m22)) / 2
    y = np.sqrt(np.maximum(0.0, 1.0 - m00 + m11 - m22)) / 2
    z = np.sqrt(np.maximum(0.0, 1.0 - m00 - m11 + m22)) / 2
    x = np.copysign(x, R[:, 2, 1] - R[:, 1, 2])
    y = np.copysign(y, R[:, 0, 2] - R[:, 2, 0])
    z = np.copysign(z, R[:, 1, 0] - R[:, 0, 1])
    q[:, 0], q[:, 1], q[:, 2], q[:, 3] = w, x, y, z
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def make_scene(P: int, seed: int, sh_coeffs: int = 16) -> Dict[str, np.ndarray]:
    """Returns float32 arrays: means3D[P,3] scales[P,2] rotations[P,4] opacities[P,1] shs[P,M,3]."""
    rng = np.random.default_rng(seed)
    hx, hy, hz = BOX / 2
    # faces: (axis, sign, area); normals point inward
    faces = [(0, +1, BOX[1] * BOX[2]), (0, -1, BOX[1] * BOX[2]), (1, +1, BOX[0] * BOX[2]),
             (1, -1, BOX[0] * BOX[2]), (2, +1, BOX[0] * BOX[1]), (2, -1, BOX[0] * BOX[1])]
    n_sphere = P // 10
    n_box = P - n_sphere
    areas = np.array([f[2] for f in faces])
    a_total = areas.sum() + 3 * 4 * math.pi * SPHERE_R ** 2
    face_of = rng.choice(6, size=n_box, p=areas / areas.sum())
    uv = rng.uniform(-1.0, 1.0, size=(n_box, 2))
    pos = np.empty((P, 3))
    nrm = np.empty((P, 3))
    half = np.array([hx, hy, hz])
    for fi, (axis, sign, _) in enumerate(faces):
        sel = np.nonzero(face_of == fi)[0]
        others = [a for a in range(3) if a != axis]
        pos[sel, axis] = sign * half[axis]
        pos[sel, others[0]] = uv[sel, 0] * half[others[0]]
        pos[sel, others[1]] = uv[sel, 1] * half[others[1]]
        nrm[sel] = 0.0
        nrm[sel, axis] = -sign
    if n_sphere:
        which = rng.integers(0, 3, size=n_sphere)
        d = _normalize(rng.normal(size=(n_sphere, 3)))
        pos[n_box:] = SPHERES[which] + SPHERE_R * d
        nrm[n_box:] = d
    nrm = _normalize(nrm + rng.normal(scale=0.1, size=(P, 3)))
    rot = _frame_quaternion(nrm, rng.uniform(0.0, 2 * math.pi, size=P))
    scales = 1.0 * math.sqrt(a_total / P) * np.exp(rng.normal(scale=0.3, size=(P, 2)))
    opac = 1.0 / (1.0 + np.exp(-rng.normal(loc=1.0, scale=1.5, size=(P, 1))))
    rgb = rng.uniform(0.0, 1.0, size=(P, 3))
    shs = rng.normal(scale=0.05, size=(P, sh_coeffs, 3))
    shs[:, 0, :] = (rgb - 0.5) / SH_C0
    perm = rng.permutation(P)  # no spatial order in memory, like a trained model after densification
    f32 = lambda a: np.ascontiguousarray(a[perm], dtype=np.float32)
    return dict(means3D=f32(pos), scales=f32(scales), rotations=f32(rot), opacities=f32(opac), shs=f32(shs))


@dataclass
class SyntheticCamera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] float32, transposed world-to-camera
    projmatrix: np.ndarray   # [4,4] float32, transposed full projection
    campos: np.ndarray       # [3] float32
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def FoVx(self) -> float:
        return 2 * math.atan(self.tanfovx)

    @property
    def FoVy(self) -> float:
        return 2 * math.atan(self.tanfovy)


def look_at_camera(eye, target, W: int, H: int, fovx_deg: float = 60.0, znear: float = 0.01,
                   zfar: float = 100.0) -> SyntheticCamera:
    eye = np.asarray(eye, dtype=np.float64)
    fwd = _normalize(np.asarray(target, dtype=np.float64) - eye)
    down = np.array([0.0, -1.0, 0.0])
    right = _normalize(np.cross(down, fwd))
    down2 = np.cross(fwd, right)
    w2c = np.eye(4)
    w2c[:3, :3] = np.stack([right, down2, fwd], axis=0)
    w2c[:3, 3] = -w2c[:3, :3] @ eye
    w2c = np.float32(w2c)
    tanx = math.tan(math.radians(fovx_deg) / 2)
    tany = tanx * H / W
    proj = np.zeros((4, 4), dtype=np.float32)
    proj[0, 0] = 1.0 / tanx
    proj[1, 1] = 1.0 / tany
    proj[3, 2] = 1.0
    proj[2, 2] = zfar / (zfar - znear)
    proj[2, 3] = -(zfar * znear) / (zfar - znear)
    view_t = np.ascontiguousarray(w2c.T)
    full_t = np.ascontiguousarray((view_t @ proj.T).astype(np.float32))
    campos = np.float32(np.linalg.inv(view_t.astype(np.float64))[3, :3])
    return SyntheticCamera(W, H, tanx, tany, view_t, full_t, campos, znear, zfar)


def make_cameras(count: int, W: int, H: int, fovx_deg: float = 60.0) -> List[SyntheticCamera]:
    """k-th of `count` cameras: on a circle r = 2.5 m (0.5 m from the walls), eye height wobbling
    around +0.2 m, looking across the room at a point 1 m beyond the centre, slightly downward."""
    cams = []
    for k in range(count):
        ang = 2 * math.pi * k / max(count, 1) + 0.3
        eye = np.array([2.5 * math.cos(ang), 0.2 + 0.15 * math.sin(3 * ang), 2.5 * math.sin(ang)])
        target = np.array([-1.0 * math.cos(ang), -0.4, -1.0 * math.sin(ang)])
        cams.append(look_at_camera(eye, target, W, H, fovx_deg))
    return cams


def make_upstream_grads(W: int, H: int, seed: int):
    """dL/dcolor [3,H,W] and dL/dallmap [7,H,W]: all channels non-zero so every backward branch runs."""
    rng = np.random.default_rng(seed + 1)
    n = float(W * H)
    return (np.float32(rng.normal(size=(3, H, W)) / n), np.float32(rng.normal(size=(7, H, W)) / n))


def to_torch(d: Dict[str, np.ndarray], device):
    import torch
    return {k: torch.from_numpy(v).to(device) for k, v in d.items()}
