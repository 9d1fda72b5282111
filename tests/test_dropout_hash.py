"""Statistical quality of the counter-based dropout hash of the LayerNorm sites (csrc/common.cuh: drop_key / drop_bits2 --
one 32-bit hash per pair of consecutive elements, two 16-bit uniforms), restated in numpy.  CPU only: the device-side
consistency of the mask between forward and backward is tests/test_kernels_gpu.py and tests/test_gemm_ln_gpu.py.
The reference draws its masks from torch's Philox stream (nn.Dropout in speechbrain's layers), which cannot be matched
bit for bit; what must hold is what training needs: the drop rate, independence between neighbouring elements, rows,
the two halves of a hash and different seeds."""
import numpy as np
import pytest

M64 = (1 << 64) - 1


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def drop_bits2(seed, pair):
    """common.cuh:drop_bits2 on a uint32 array of pair indices."""
    m = mix64(seed)
    s0, s1 = np.uint32(m & 0xFFFFFFFF), np.uint32(m >> 32)
    with np.errstate(over="ignore"):
        x = pair.astype(np.uint32) * np.uint32(0x9E3779B1) + s0
        x ^= x >> np.uint32(16)
        x = x * np.uint32(0x21F0AAAD) + s1
        x ^= x >> np.uint32(15)
        x = x * np.uint32(0x735A2D97)
        x ^= x >> np.uint32(15)
    return x


def keep_mask(seed, rows, C, p):
    e = np.arange(rows * C, dtype=np.uint64)
    h = drop_bits2(seed, (e >> np.uint64(1)).astype(np.uint32))
    u = np.where(e & np.uint64(1), h >> np.uint32(16), h & np.uint32(0xFFFF))
    return (u >= np.uint32(int(np.float32(p) * np.float32(65536.0)))).reshape(rows, C)


def corr(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.corrcoef(a, b)[0, 1])


@pytest.mark.parametrize("p", [0.1, 0.2, 0.5])
def test_drop_rate_and_independence(p):
    rows, C = 2048, 384
    k = keep_mask(0x1234ABCD5678, rows, C, p)
    n = k.size
    sigma = (p * (1 - p) / n) ** 0.5
    assert abs((1 - k.mean()) - p) < 5 * sigma + 1.0 / 65536            # rate (threshold quantised to 1 / 65536)
    bound = 5 / n ** 0.5                                                # five sigma of a sample correlation of independent bits
    assert abs(corr(k[:, 0::2], k[:, 1::2])) < bound                    # the two halves of one hash
    assert abs(corr(k[:, 1:-1:2], k[:, 2::2])) < bound                  # neighbours from different hashes
    assert abs(corr(k[:-1], k[1:])) < bound                             # the same column of neighbouring rows
    assert abs(corr(k[:, :-4], k[:, 4:])) < bound                       # one float4 group apart
    k2 = keep_mask(0x1234ABCD5679, rows, C, p)                          # the next seed (next site / next step)
    assert abs(corr(k, k2)) < bound
    # per-row and per-column drop counts are binomial: no row or column is systematically favoured
    for counts, m in (((~k).sum(1), C), ((~k).sum(0), rows)):
        z = (counts - m * p) / (m * p * (1 - p)) ** 0.5
        assert abs(z.mean()) < 5 / len(z) ** 0.5 and 0.8 < z.std() < 1.2 and np.abs(z).max() < 6


def test_uniforms_are_flat():
    h = drop_bits2(77, np.arange(1 << 20, dtype=np.uint32))
    for u in (h & np.uint32(0xFFFF), h >> np.uint32(16)):
        hist = np.bincount((u >> np.uint32(8)).astype(np.int64), minlength=256)      # 256 buckets of the 16-bit uniform
        expect = len(u) / 256
        chi2 = float(((hist - expect) ** 2 / expect).sum())
        assert chi2 < 255 + 6 * (2 * 255) ** 0.5                                      # chi-square, 255 dof, +6 sigma
