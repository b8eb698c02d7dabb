"""CPU restatement of the dropout mask of ``opn_dropout`` (include/opnet_b200.h).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as oracle/opnet_oracle.py).

The reference gets its train-mode dropout from PyTorch (``nn.TransformerEncoderLayer(dropout=0.1)``,
baselines/learned_models.py:166), whose random stream is an implementation detail of the installed
PyTorch build and cannot be reproduced by another implementation.  What can be pinned is the
generator this library specifies: Philox4x32-10 (Salmon et al., "Parallel Random Numbers: As Easy as
1, 2, 3", SC'11), counter = (offset + i // 4, 0), key = seed, element ``i`` takes word ``i % 4`` and
is kept when that word is ``>= p * 2**32``.  ``tests/test_oracle_golden.py`` checks this restatement
against the Random123 known-answer vectors; the GPU tests check the kernel against it bit for bit.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """counter uint32 [n, 4], key (k0, k1) -> uint32 [n, 4]."""
    c = [counter[:, j].astype(np.uint64) for j in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _LO, p1 >> np.uint64(32), p1 & _LO
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def keep_mask(n: int, p: float, seed: int, offset: int) -> np.ndarray:
    """bool [n]: which elements opn_dropout(n, ..., p, seed, offset) keeps."""
    blocks = (n + 3) // 4
    ctr = np.uint64(offset) + np.arange(blocks, dtype=np.uint64)
    counter = np.zeros((blocks, 4), dtype=np.uint32)
    counter[:, 0] = (ctr & _LO).astype(np.uint32)
    counter[:, 1] = (ctr >> np.uint64(32)).astype(np.uint32)
    words = philox4x32_10(counter, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n]
    t = float(np.float32(p)) * 4294967296.0
    threshold = 4294967295 if t >= 4294967295.0 else int(t)
    return words >= np.uint32(threshold)


def dropout(x: np.ndarray, p: float, seed: int, offset: int) -> np.ndarray:
    """float32 result of opn_dropout on a flat float32 array."""
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep_mask(x.size, p, seed, offset), x.reshape(-1).astype(np.float32) * scale,
                    np.float32(0.0)).reshape(x.shape)
