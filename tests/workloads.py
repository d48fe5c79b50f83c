"""Synthetic point clouds shared by tests, golden generation and bench.py."""
import numpy as np


def make_cloud(kind, n, seed=0):
    g = np.random.default_rng(seed)
    if kind == "sphere":                      # C1: README quick-start shape
        X = g.normal(size=(n, 3))
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        return X
    if kind == "torus":                       # C2 / C4
        u, v = g.uniform(0, 2 * np.pi, (2, n))
        R, r = 1.0, 0.35
        return np.stack([(R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)], 1)
    if kind == "scalp":                       # C3: upper hemisphere z >= 0.1
        out = []
        while sum(len(o) for o in out) < n:
            X = g.normal(size=(2 * n, 3))
            X /= np.linalg.norm(X, axis=1, keepdims=True)
            out.append(X[X[:, 2] >= 0.1])
        return np.concatenate(out)[:n]
    if kind == "moebius":                     # NON-orientable 2-manifold: no consistent gauge orientation exists
        u = g.uniform(0, 2 * np.pi, n)
        v = g.uniform(-0.4, 0.4, n)
        return np.stack([(1 + 0.5 * v * np.cos(u / 2)) * np.cos(u), (1 + 0.5 * v * np.cos(u / 2)) * np.sin(u),
                         0.5 * v * np.sin(u / 2)], 1)
    if kind == "flat3torus":                  # 3-manifold in R^6 (d=3 blocks)
        a = g.uniform(0, 2 * np.pi, (n, 3))
        return np.concatenate([np.cos(a), np.sin(a)], 1)[:, [0, 3, 1, 4, 2, 5]].copy()
    if kind == "sheet_R20":                   # 2-manifold in R^20 (D > 15 -> sklearn brute kNN)
        u, v = g.uniform(0, 1, (2, n))
        cols = [u, v]
        for f in range(1, 10):
            cols.append(0.1 * np.sin(f * u + 0.3 * f * v))
            cols.append(0.1 * np.cos(f * v - 0.2 * f * u))
        X = np.stack(cols, 1)
        X += g.normal(scale=2e-3, size=X.shape)   # full local rank (pyx:426-428)
        return X
    if kind == "manifold5_R32":               # C5
        th = g.uniform(0, 2 * np.pi, (n, 5))
        cols = [np.cos(th), np.sin(th)]
        extra = []
        for i in range(5):
            extra.append(0.05 * np.cos(2 * th[:, i]))
            extra.append(0.05 * np.sin(2 * th[:, i]))
        pairs = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0), (0, 2)]
        for (i, j) in pairs:
            extra.append(0.05 * np.cos(th[:, i] + th[:, j]))
            extra.append(0.05 * np.sin(th[:, i] + th[:, j]))
        X = np.concatenate(cols + [np.stack(extra, 1)], 1)
        assert X.shape[1] == 32
        X += np.random.default_rng(seed + 1).normal(scale=1e-3, size=X.shape)
        return X
    raise ValueError(kind)
