"""numpy restatement of the kernels' dropout mask (c2v_common.cuh: philox4x32_10 / dropout_*):
lets the tests hand the CPU oracle the exact mask the CUDA path drew, so training-mode
forward/backward parity is checked exactly, not statistically."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0; p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask; k1 = (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def dropout_mask(seed, n_rows, H, p):
    """multiplicative mask [n_rows, H]: 0 or 1/(1-p); element (row, col) uses word col&3 of the
    Philox block with counter (row_lo, row_hi, col>>2, 0) and key (seed_lo, seed_hi)."""
    rows = np.arange(n_rows, dtype=np.uint64)[:, None]
    c4 = np.arange((H + 3) // 4, dtype=np.uint64)[None, :]
    r0 = np.broadcast_to(rows & np.uint64(0xFFFFFFFF), (n_rows, c4.shape[1]))
    r1 = np.broadcast_to(rows >> np.uint64(32), (n_rows, c4.shape[1]))
    cc = np.broadcast_to(c4, (n_rows, c4.shape[1]))
    w = philox4x32_10(r0, r1, cc, np.zeros_like(cc), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(w, axis=-1).reshape(n_rows, -1)[:, :H]
    u = (bits >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u >= np.float32(p), scale, np.float32(0.0)).astype(np.float32)
