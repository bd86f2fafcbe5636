"""TEST INFRASTRUCTURE (oracle): the level-set distance functions and closest-point maps of
StefanDG/useful_routines.jl:26-74, restated (inputs: coords [2, npts])."""
import numpy as np


def plane_distance_function(coords, normal, x0):
    """useful_routines.jl:26-28"""
    return (np.asarray(coords) - np.asarray(x0).reshape(2, 1)).T @ np.asarray(normal)


def closest_point_on_plane(querypoints, normal, x0):
    """useful_routines.jl:30-34"""
    q, n = np.asarray(querypoints, dtype=float), np.asarray(normal, dtype=float)
    normalcomp = n @ (q - np.asarray(x0).reshape(2, 1))
    return q - np.outer(n, normalcomp)


def circle_distance_function(coords, center, radius):
    """useful_routines.jl:36-40: positive inside"""
    d = np.asarray(coords, dtype=float) - np.asarray(center).reshape(2, 1)
    return radius - np.sqrt(np.sum(d ** 2, axis=0))


def closest_point_on_arc(querypoints, center, radius):
    """useful_routines.jl:42-48"""
    rel = np.asarray(querypoints, dtype=float) - np.asarray(center).reshape(2, 1)
    theta = np.angle(rel[0] + 1j * rel[1])
    return np.asarray(center).reshape(2, 1) + radius * np.vstack([np.cos(theta), np.sin(theta)])


def corner_distance_function(points, xc):
    """useful_routines.jl:50-68 (vector and matrix methods)"""
    pts = np.asarray(points, dtype=float)
    xc = np.asarray(xc, dtype=float)
    if pts.ndim == 1:
        v = xc - pts
        if np.all(pts <= xc):
            return v.min()
        if np.all(pts > xc):
            return -np.sqrt(v @ v)
        return v[1] if pts[1] > xc[1] else v[0]
    return np.array([corner_distance_function(pts[:, k], xc) for k in range(pts.shape[1])])
