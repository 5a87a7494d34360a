"""TEST INFRASTRUCTURE — numpy restatement of the Philox4x32 stream the x2k kernels use for dropout
(x2vlm_b200/csrc/common.cuh: philox4x32<7> / drop8), to build the exact keep-mask on the CPU.

Published algorithm: Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11), Philox-4x32 with
10 rounds, multipliers 0xD2511F53 / 0xCD9E8D57, Weyl keys 0x9E3779B9 / 0xBB67AE85.  The kernels' counter is
(ctr_lo, ctr_hi, 0x2B992DDF, 0); the known-answer test in tests/test_oracle_golden.py pins the round function
against the Random123 vector for the all-zero counter/key."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_raw(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox4x32(seed, ctr):
    """Kernel convention: key = seed (lo, hi), counter = (ctr lo, ctr hi, 0x2B992DDF, 0). ctr: uint64 array."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    return philox4x32_raw(ctr & MASK, ctr >> np.uint64(32), np.full_like(ctr, 0x2B992DDF), np.zeros_like(ctr),
                          int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)


def keep_scale(seed, offset, n_elems, p, rounds=7):
    """float32 [n_elems] keep-scales of the kernels' dropout stream (common.cuh: drop8): element e draws 16 bits from
    Philox4x32-7(seed, counter = offset + (e >> 3)), word (e & 7) >> 1, low half for even e / high half for odd e; it is
    kept iff r16 >= thr = round(p * 65536) and then scaled by 65536 / (65536 - thr)."""
    n8 = (n_elems + 7) // 8
    ctr = np.uint64(offset) + np.arange(n8, dtype=np.uint64)
    words = philox4x32_raw(ctr & MASK, ctr >> np.uint64(32), np.full_like(ctr, 0x2B992DDF), np.zeros_like(ctr),
                           int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF, rounds=rounds)
    w = np.stack(words, axis=1)  # [n8, 4]
    r16 = np.stack([w & np.uint64(0xFFFF), w >> np.uint64(16)], axis=2).reshape(-1)[:n_elems]
    thr = int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)) if p > 0 else 0
    scale = np.float32(65536.0) / (np.float32(65536.0) - np.float32(thr))
    return np.where(r16 >= np.uint64(thr), scale, np.float32(0.0)).astype(np.float32)
