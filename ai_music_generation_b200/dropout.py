"""Host-side twin of csrc/dropout.cuh: site keys and keep-masks as numpy arrays (uint32 arithmetic), used to derive the
per-site keys the kernels receive and — in tests — to feed the CPU oracle the exact masks the kernels applied."""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _fmix32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & M32
    x ^= x >> np.uint64(16)
    return x


def site_key(seed: int, site: int) -> int:
    """site 0 = embedding; block l: 1+3l attention probabilities, 2+3l attention residual, 3+3l MLP residual."""
    return int(_fmix32(np.uint64((seed ^ ((site * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF)))


def threshold(p: float) -> int:
    return 0 if p <= 0 else min(65535, int(p * 65536.0 + 0.5))


def keep_mask(key: int, rows: int, cols: int, p: float, row_offset: int = 0) -> np.ndarray:
    """bool [rows, cols]; row r uses counter row_offset + r."""
    thr = np.uint64(threshold(p))
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row_offset)) & M32
    row_key = _fmix32((np.uint64(key) + r * np.uint64(0x85EBCA77)) & M32)              # [rows]
    pair = (np.arange((cols + 1) // 2, dtype=np.uint64) * np.uint64(0x27D4EB2F)) & M32     # [pairs]
    bits = _fmix32(row_key[:, None] ^ pair[None, :])                                      # [rows, pairs]
    lo = (bits & np.uint64(0xFFFF)) >= thr
    hi = (bits >> np.uint64(16)) >= thr
    out = np.empty((rows, 2 * pair.shape[0]), dtype=bool)
    out[:, 0::2] = lo
    out[:, 1::2] = hi
    return out[:, :cols]


def attention_keep_mask(key: int, B: int, H: int, T: int, p: float) -> np.ndarray:
    """bool [B, H, T, T]: row counter (b*H + h)*T + q, column = key position.  The attention-probability site uses the
    cheaper generator of csrc/dropout.cuh (Weyl step + folded 32x32->64 multiply, two 15-bit lanes per word)."""
    rows = B * H * T
    thr = np.uint64(0 if p <= 0 else min(32767, int(p * 32768.0 + 0.5)))
    r = np.arange(rows, dtype=np.uint64) & M32
    a = _fmix32((np.uint64(key) + r * np.uint64(0x85EBCA77)) & M32)                       # [rows] (= drop_row_key)
    b = (a * np.uint64(0x9E3779B1) + np.uint64(0x7F4A7C15)) & M32
    pair = (np.arange((T + 1) // 2, dtype=np.uint64) * np.uint64(0x53C5CA59)) & M32
    s = (a[:, None] + pair[None, :]) & M32
    c = (s ^ b[:, None]) * s                                                              # exact: both factors < 2^32
    h = ((c & M32) ^ (c >> np.uint64(32))) & M32
    out = np.empty((rows, 2 * pair.shape[0]), dtype=bool)
    out[:, 0::2] = (h & np.uint64(0x7FFF)) >= thr
    out[:, 1::2] = ((h >> np.uint64(16)) & np.uint64(0x7FFF)) >= thr
    return out[:, :T].reshape(B, H, T, T)
