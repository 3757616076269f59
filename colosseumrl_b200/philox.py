"""Host-side Philox4x32-10, bit-identical to the device generator in csrc/philox.cuh.

The random policy used for parity tests and for the benchmark is *stateless*:
the four 32-bit words for (environment e, step t) are

    philox4x32_10(counter=(e_lo, e_hi, t, tag), key=(seed_lo, seed_hi))

so an environment's action stream does not depend on batch size, on how the batch
is sharded over GPUs, or on episode boundaries.  This is
the Random123 generator; numpy's own ``Philox`` is the 4x64 variant and is NOT
compatible.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

# stream tags (counter word 3): one per consumer so streams never collide
TAG_TRON = 1
TAG_BLOKUS = 2
TAG_TTT = 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable arrays of 32-bit values.

    Returns four uint32 arrays.
    """
    c0, c1, c2, c3, k0, k1 = (np.asarray(v).astype(np.uint64) & _MASK for v in (c0, c1, c2, c3, k0, k1))
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> _S32) ^ c1 ^ k0
        n1 = p1 & _MASK
        n2 = (p0 >> _S32) ^ c3 ^ k1
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & _MASK
        k1 = (k1 + W1) & _MASK
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def env_step_words(seed: int, env_ids, step: int, tag: int) -> np.ndarray:
    """Random words for a batch of environments at one step: uint32 [B, 4]."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(env_ids & _MASK, env_ids >> _S32, np.uint64(step & 0xFFFFFFFF), np.uint64(tag),
                      np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32))
    return np.stack(r, axis=-1)
