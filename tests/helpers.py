"""Shared synthetic-data helpers for the tests (SURVEY.md §8d distribution)."""
import numpy as np


def clustered(n, dim, seed, n_clusters=64, dup_frac=0.01):
    """Unit rows around `n_clusters` centres (+ a few exact duplicates to force score ties)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_clusters, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    noise = rng.standard_normal((n, dim)).astype(np.float32)
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    x = 0.6 * centres[np.arange(n) % n_clusters] + 0.8 * noise
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    ndup = int(n * dup_frac)
    if ndup:
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        x[dst] = x[src]
    return x.astype(np.float32), centres


def queries_for(centres, x, nq, seed):
    rng = np.random.default_rng(seed)
    dim = centres.shape[1]
    noise = rng.standard_normal((nq, dim)).astype(np.float32)
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    q = 0.6 * centres[rng.integers(0, len(centres), nq)] + 0.8 * noise
    if nq >= 4:
        q[0] = x[rng.integers(0, len(x))]          # exact corpus row
        q[1] = 7.25 * q[1]                          # un-normalised
    return q.astype(np.float32)
