"""Seeded synthetic inputs shared by the tests, smoke() and bench.py."""
import numpy as np


def uniform(n, d, seed=1):
    """Program.cs:163-181 shape: x = U[0,1)*2-1 float32, ids 0..n-1."""
    rng = np.random.default_rng(seed)
    rows = (rng.random((n, d), dtype=np.float32) * np.float32(2) - np.float32(1)).astype(np.float32)
    return np.arange(n, dtype=np.int64), rows


def unit_gaussian(n, d, seed=2, chunk=1 << 18):
    """deep-image-96-angular shape: N(0,1) rows, L2-normalised, float32, ids 0..n-1."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, d), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        g = rng.standard_normal((e - s, d), dtype=np.float32)
        g /= np.linalg.norm(g, axis=1, keepdims=True).astype(np.float32)
        rows[s:e] = g
    return np.arange(n, dtype=np.int64), rows


def one_hot(d=1536):
    """Program.cs:54-66 crafted set: d vectors e_i with id i."""
    return np.arange(d, dtype=np.int64), np.eye(d, dtype=np.float32)


def grid2d(m):
    """MemoryVectorIndexTests.cs:10-92 style m x m grid normalised into [-1, 1]."""
    xs = np.linspace(-1, 1, m, dtype=np.float32)
    g = np.stack(np.meshgrid(xs, xs, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    return np.arange(g.shape[0], dtype=np.int64), g
