"""
Host-side mirror of the reference's start-up camera
(/root/reference/src/utils/camera_controller.cpp:30-41,77-80): the benchmark camera
of every BASELINE config.  The fly-camera controls themselves are out of
scope (windowing).
"""
import math

import numpy as np

from .layouts import CAMERA_DT


def default_camera(width: int, height: int, position=(0.0, -1.0, 1.0), aperture=0.0, focus_distance=10.0,
                   yaw=1.570796327, pitch=1.570796327) -> np.ndarray:
    """yaw = pitch = MATH_PIDIV2, world up (0,0,1), fov = 75*3.1415/180, aspect = w/h; every step in
    float32 with correctly rounded sin/cos/sqrt (camera_controller.cpp:77-80, mathlib.hpp:47-48)."""
    f = np.float32
    yaw, pitch = f(yaw), f(pitch)
    cs = lambda a: f(math.cos(float(a)))
    sn = lambda a: f(math.sin(float(a)))
    front = np.array([cs(yaw) * sn(pitch), sn(yaw) * sn(pitch), cs(pitch)], dtype=np.float32)
    up_w = np.array([0, 0, 1], dtype=np.float32)

    def cross(a, b):
        return np.array([f(a[1] * b[2]) - f(a[2] * b[1]), f(a[2] * b[0]) - f(a[0] * b[2]),
                         f(a[0] * b[1]) - f(a[1] * b[0])], dtype=np.float32)
    r = cross(front, up_w)
    ln = f(math.sqrt(float(f(f(f(r[0] * r[0]) + f(r[1] * r[1])) + f(r[2] * r[2])))))
    right = np.array([r[0] / ln, r[1] / ln, r[2] / ln], dtype=np.float32)
    up = cross(right, front)
    cam = np.zeros((), dtype=CAMERA_DT)
    cam["position"][:3] = position
    cam["front"][:3] = front
    cam["up"][:3] = up
    cam["fov"] = f(f(f(75.0) * f(3.1415)) / f(180.0))
    cam["aspect_ratio"] = f(width) / f(height)
    cam["aperture"] = aperture
    cam["focus_distance"] = focus_distance
    return cam
