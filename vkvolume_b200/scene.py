"""Host-side helpers around the C ABI: cameras, image transforms, small synthetic volumes.

Pure parameter plumbing (no voxel arithmetic of the hot path lives here).
"""
from __future__ import annotations

import math

import numpy as np

from .capi import CameraDesc


def quat_from_matrix(R: np.ndarray):
    """Rotation matrix (3x3, columns = basis vectors) -> quaternion (x, y, z, w)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w = (R[2, 1] - R[1, 2]) / s
        x = 0.25 * s
        y = (R[0, 1] + R[1, 0]) / s
        z = (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w = (R[0, 2] - R[2, 0]) / s
        x = (R[0, 1] + R[1, 0]) / s
        y = 0.25 * s
        z = (R[1, 2] + R[2, 1]) / s
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w = (R[1, 0] - R[0, 1]) / s
        x = (R[0, 2] + R[2, 0]) / s
        y = (R[1, 2] + R[2, 1]) / s
        z = 0.25 * s
    return (x, y, z, w)


def look_at_camera(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), yfov=1.0, aspect=1.0, znear=1.0, zfar=4000.0,
                   node_scale=(100.0, 100.0, 100.0)) -> CameraDesc:
    """Camera node (translation + rotation) looking from `eye` to `target`; -Z is forward as in glTF."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    back = eye - target
    back /= np.linalg.norm(back)
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    up2 = np.cross(back, right)
    R = np.stack([right, up2, back], axis=1)
    return CameraDesc(translation=tuple(eye), rotation=quat_from_matrix(R), yfov=yfov, aspect=aspect, znear=znear,
                      zfar=zfar, node_scale=node_scale)


def image_transform(voxel_size, extent, axis_angle=(1.0, 0.0, 0.0, 0.0)):
    """rotate(radians(angle), axis) * scale(voxel_size * extent), column-major list of 16 (load_volume.cpp:82-83)."""
    ax = np.asarray(axis_angle[:3], dtype=np.float64)
    ang = math.radians(float(np.float32(axis_angle[3])))
    phys = [float(np.float32(v) * np.float32(e)) for v, e in zip(voxel_size, extent)]
    n = np.linalg.norm(ax)
    ax = ax / n if n > 0 else ax
    c, s = math.cos(ang), math.sin(ang)
    t = (1 - c) * ax
    R = np.eye(4)
    R[0, 0] = c + t[0] * ax[0]; R[1, 0] = t[0] * ax[1] + s * ax[2]; R[2, 0] = t[0] * ax[2] - s * ax[1]
    R[0, 1] = t[1] * ax[0] - s * ax[2]; R[1, 1] = c + t[1] * ax[1]; R[2, 1] = t[1] * ax[2] + s * ax[0]
    R[0, 2] = t[2] * ax[0] + s * ax[1]; R[1, 2] = t[2] * ax[1] - s * ax[0]; R[2, 2] = c + t[2] * ax[2]
    S = np.diag([phys[0], phys[1], phys[2], 1.0])
    M = R @ S
    return [float(np.float32(M[r, c_])) for c_ in range(4) for r in range(4)]        # column-major


def blobs_volume(shape_dhw, seed=0, n_blobs=12, noise=4) -> np.ndarray:
    """Small seeded test volume [z, y, x]: Gaussian blobs + uniform noise (numpy; for tests)."""
    rng = np.random.default_rng(seed)
    D, H, W = shape_dhw
    z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    v = np.zeros(shape_dhw, dtype=np.float32)
    m = min(shape_dhw)
    for _ in range(n_blobs):
        c = rng.uniform(0, 1, 3) * np.array([D, H, W])
        s = rng.uniform(0.04, 0.14) * m
        a = rng.uniform(64, 255)
        v += a * np.exp(-0.5 * ((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2) / (s * s))
    v += rng.integers(0, noise + 1, size=shape_dhw)
    return np.clip(v, 0, 255).astype(np.uint8)
