"""CPU check of the algebra behind tc_pair16_kernel (parrot_tts_b200/csrc/tc_pair.cu): a K-tap convolution evaluated with G output
positions per accumulator row against a block-Toeplitz operand equals the tap-by-tap convolution, and every output receives its taps
in the same order (the zero blocks only add zeros) - which is why the two CUDA kernels produce the same bits.  Also the index map the
kernel uses to expand the ordinary packed conv2 weights (tc_pack_conv_weight layout, N = 16) into that operand."""
import numpy as np
import pytest

G = 4


def _toeplitz(W):
    """W [K][ci][co] -> T [G + K - 1][ci][G * co]: T[s][:, g*C:(g+1)*C] = W[s - g] where 0 <= s - g < K (tc_pair.cu, kernel prologue)."""
    K, C, _ = W.shape
    T = np.zeros((G + K - 1, C, G * C), W.dtype)
    for s in range(G + K - 1):
        for g in range(G):
            if 0 <= s - g < K:
                T[s][:, g * C:(g + 1) * C] = W[s - g]
    return T


@pytest.mark.parametrize("K", [3, 7, 11])
def test_time_grouped_conv_equals_tap_by_tap(K):
    rng = np.random.default_rng(K)
    C, rows = 16, 12
    # (integer-valued data: every sum is exact, so the comparison does not depend on the BLAS's summation order inside a matmul)
    xt = rng.integers(-64, 64, (G * rows + K - 1, C)).astype(np.float64)    # xt[p] = conv1 output at position o0 - (K-1)/2 + p
    W = rng.integers(-64, 64, (K, C, C)).astype(np.float64)
    # tap by tap (tc_pair_kernel): y[t] = sum_j xt[t + j] . W[j], taps in increasing j
    ref = np.zeros((G * rows, C), np.float64)
    for j in range(K):
        ref += xt[j:j + G * rows] @ W[j]
    # time-grouped: accumulator row r = positions 4r .. 4r+3; k-step s reads position 4r + s of every row
    T = _toeplitz(W)
    acc = np.zeros((rows, G * C), np.float64)
    for s in range(G + K - 1):
        acc += xt[s:s + G * rows:G] @ T[s]
    got = acc.reshape(rows * G, C)
    assert np.array_equal(got, ref)          # same products per output (in the same tap order), exact zeros elsewhere


def test_rows_a_super_tile_can_store():
    # 4 conv1 M tiles = 512 xt positions; output (r, g) needs xt up to 4r + g + K - 1 <= 511  ->  rows_out = (512 - (K - 1)) // 4
    for K, rows_out in ((3, 127), (7, 126), (11, 125), (15, 124)):
        assert (512 - (K - 1)) // 4 == rows_out
        assert 4 * (rows_out - 1) + 3 + K - 1 <= 511 < 4 * rows_out + 3 + K - 1 + 4


def test_expansion_index_map_from_the_packed_weights():
    """Packed conv weight block of one tap (N = 16, no CTA pair): [K half h][plane: hi rows 0-15 | lo rows 16-31][16 bytes = 8 ci].  The
    kernel copies 16-byte vectors: dst (s, h, n) <- src ((s - g) * 2 + h) * 32 + (n >= 64) * 16 + co with g = (n % 64) // 16, co = n % 16."""
    K, C = 7, 16
    rng = np.random.default_rng(0)
    w_hi = rng.integers(1, 1 << 15, (K, C, C)).astype(np.int64)      # [tap][ci][co] stand-ins for the hi / lo halves
    w_lo = rng.integers(1, 1 << 15, (K, C, C)).astype(np.int64)
    src = np.zeros((K * 2 * 32, 8), np.int64)                           # vectors of 8 input channels
    for j in range(K):
        for h in range(2):
            for co in range(C):
                src[(j * 2 + h) * 32 + co] = w_hi[j, 8 * h:8 * h + 8, co]
                src[(j * 2 + h) * 32 + 16 + co] = w_lo[j, 8 * h:8 * h + 8, co]
    dst = np.zeros(((G + K - 1) * 256, 8), np.int64)
    for i in range((G + K - 1) * 256):
        s, h, n = i >> 8, (i >> 7) & 1, i & 127
        g, co, j = (n & 63) >> 4, n & 15, (i >> 8) - ((n & 63) >> 4)
        if 0 <= j < K:
            dst[i] = src[(j * 2 + h) * 32 + (n >> 6) * 16 + co]
    T_hi, T_lo = _toeplitz(w_hi), _toeplitz(w_lo)
    for s in range(G + K - 1):
        for h in range(2):
            for n in range(128):
                want = (T_hi if n < 64 else T_lo)[s][8 * h:8 * h + 8, n % 64]
                assert np.array_equal(dst[s * 256 + h * 128 + n], want), (s, h, n)
