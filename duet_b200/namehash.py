"""128-bit read-name hash used to key the support-read join.

The reference joins support reads to haplotags on the QNAME *string*
(/root/reference/src/duet/sv_phasing_fn.py:29,47-48).  The device joins on a
64-bit key (``lo``) and verifies the second 64 bits (``hi``) whenever two keys
compare equal, so a 64-bit collision between different names is detected and
reported instead of silently mis-joining.

The function is a murmur3-x64-128 style mixer defined here (it only has to be
self-consistent between this file and ``csrc/decode.cpp::duet_hash128``):

  * little-endian 16-byte blocks (k1, k2) are mixed into (h1, h2);
  * the trailing ``n % 16`` bytes are zero padded and mixed without the
    inter-lane step;
  * ``lo == 0xFFFF_FFFF_FFFF_FFFF`` is the table's EMPTY sentinel, so ``lo`` is
    remapped to ``0xFFFF_FFFF_FFFF_FFFE`` in that case (the ``hi`` check still
    separates the two names).

Three implementations, all bit-identical: ``hash128`` (pure Python, the
definition), ``hash128_fixed`` (numpy, vectorised over equal-length names) and
the C++ one behind ``duet_hash_names``.
"""
from __future__ import annotations

import numpy as np

MASK = (1 << 64) - 1
C1 = 0x87C37B91114253D5
C2 = 0x4CF5AD432745937F
SEED1 = 0x9E3779B97F4A7C15
SEED2 = 0xD1B54A32D192ED03
EMPTY_KEY = MASK


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK


def _fmix(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & MASK
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & MASK
    k ^= k >> 33
    return k


def hash128(name: bytes | str) -> tuple[int, int]:
    """Return ``(lo, hi)`` for one read name.  Pure Python definition."""
    if isinstance(name, str):
        name = name.encode("ascii")
    n = len(name)
    h1, h2 = SEED1, SEED2
    nblocks = n // 16
    for b in range(nblocks):
        k1 = int.from_bytes(name[16 * b:16 * b + 8], "little")
        k2 = int.from_bytes(name[16 * b + 8:16 * b + 16], "little")
        k1 = (k1 * C1) & MASK
        k1 = _rotl(k1, 31)
        k1 = (k1 * C2) & MASK
        h1 ^= k1
        h1 = _rotl(h1, 27)
        h1 = (h1 + h2) & MASK
        h1 = (h1 * 5 + 0x52DCE729) & MASK
        k2 = (k2 * C2) & MASK
        k2 = _rotl(k2, 33)
        k2 = (k2 * C1) & MASK
        h2 ^= k2
        h2 = _rotl(h2, 31)
        h2 = (h2 + h1) & MASK
        h2 = (h2 * 5 + 0x38495AB5) & MASK
    tail = name[16 * nblocks:] + b"\0" * 16
    k1 = int.from_bytes(tail[0:8], "little")
    k2 = int.from_bytes(tail[8:16], "little")
    k2 = (k2 * C2) & MASK
    k2 = _rotl(k2, 33)
    k2 = (k2 * C1) & MASK
    h2 ^= k2
    k1 = (k1 * C1) & MASK
    k1 = _rotl(k1, 31)
    k1 = (k1 * C2) & MASK
    h1 ^= k1
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & MASK
    h2 = (h2 + h1) & MASK
    h1 = _fmix(h1)
    h2 = _fmix(h2)
    h1 = (h1 + h2) & MASK
    h2 = (h2 + h1) & MASK
    if h1 == EMPTY_KEY:
        h1 = EMPTY_KEY - 1
    return h1, h2


def _np_rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << np.uint64(r)) | (x >> np.uint64(64 - r))


def _np_fmix(k: np.ndarray) -> np.ndarray:
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xFF51AFD7ED558CCD)
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xC4CEB9FE1A85EC53)
    k = k ^ (k >> np.uint64(33))
    return k


def hash128_fixed(names: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Vectorised ``hash128`` over an ``(N, L)`` uint8 array (or ``S<L>`` array)
    of equal-length names.  Returns ``(lo, hi)`` as uint64 arrays."""
    if names.dtype.kind == "S":
        width = names.dtype.itemsize
        names = np.frombuffer(np.ascontiguousarray(names).tobytes(), dtype=np.uint8).reshape(-1, width)
    names = np.ascontiguousarray(names, dtype=np.uint8)
    n_rows, n = names.shape
    nblocks = n // 16
    padded = np.zeros((n_rows, 16 * (nblocks + 1)), dtype=np.uint8)
    padded[:, :n] = names
    words = padded.view("<u8")  # (n_rows, 2*(nblocks+1))
    c1, c2 = np.uint64(C1), np.uint64(C2)
    h1 = np.full(n_rows, SEED1, dtype=np.uint64)
    h2 = np.full(n_rows, SEED2, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for b in range(nblocks):
            k1 = words[:, 2 * b].copy()
            k2 = words[:, 2 * b + 1].copy()
            k1 = _np_rotl(k1 * c1, 31) * c2
            h1 = h1 ^ k1
            h1 = _np_rotl(h1, 27) + h2
            h1 = h1 * np.uint64(5) + np.uint64(0x52DCE729)
            k2 = _np_rotl(k2 * c2, 33) * c1
            h2 = h2 ^ k2
            h2 = _np_rotl(h2, 31) + h1
            h2 = h2 * np.uint64(5) + np.uint64(0x38495AB5)
        k1 = words[:, 2 * nblocks].copy()
        k2 = words[:, 2 * nblocks + 1].copy()
        h2 = h2 ^ (_np_rotl(k2 * c2, 33) * c1)
        h1 = h1 ^ (_np_rotl(k1 * c1, 31) * c2)
        h1 = h1 ^ np.uint64(n)
        h2 = h2 ^ np.uint64(n)
        h1 = h1 + h2
        h2 = h2 + h1
        h1 = _np_fmix(h1)
        h2 = _np_fmix(h2)
        h1 = h1 + h2
        h2 = h2 + h1
    h1[h1 == np.uint64(EMPTY_KEY)] = np.uint64(EMPTY_KEY - 1)
    return h1, h2


def hash_names(names) -> tuple[np.ndarray, np.ndarray]:
    """Hash an iterable of str/bytes names (any lengths).  Pure Python loop;
    the C++ decoder (``duet_b200.decode``) is the fast path."""
    names = list(names)
    lo = np.empty(len(names), dtype=np.uint64)
    hi = np.empty(len(names), dtype=np.uint64)
    for i, nm in enumerate(names):
        a, b = hash128(nm)
        lo[i] = a
        hi[i] = b
    return lo, hi
