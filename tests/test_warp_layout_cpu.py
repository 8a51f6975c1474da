"""Lane-level emulation (numpy, CPU) of the register layouts attn_warp.cu builds its mma.sync.m16n8k16 fragments from.

The kernel never transposes or shuffles operands: it relies on two identities that are checked here against plain
matrix products, using the fragment ownership rules of the PTX ISA (g = lane >> 2, tig = lane & 3):
  * "R layout" (rows g, g+8; columns {8 tig..} and {32 + 8 tig..}): both operands of X . Y^T use the SAME permutation of
    the 64-wide contraction index, so register 2 kk / 2 kk + 1 of a row serve k-step kk directly;
  * "P layout" (rows 2 tig, 2 tig+1, 2 tig+8, 2 tig+9; columns 8 g..8 g+7): for A[16x16] . X[16x64] the B fragment of
    n-tile nt is element nt of the four row vectors, and accumulator (nt, j) of row g is OUTPUT column 16 tig + 8 j + nt.
"""
import numpy as np


def _mma_m16n8k16(c, a_frags, b_frags):
    """One warp-wide mma: a_frags[lane] = (a0, a1, a2, a3), b_frags[lane] = (b0, b1), each a pair of values."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        a0, a1, a2, a3 = a_frags[lane]
        b0, b1 = b_frags[lane]
        A[g, 2 * t:2 * t + 2] = a0
        A[g + 8, 2 * t:2 * t + 2] = a1
        A[g, 2 * t + 8:2 * t + 10] = a2
        A[g + 8, 2 * t + 8:2 * t + 10] = a3
        B[2 * t:2 * t + 2, g] = b0
        B[2 * t + 8:2 * t + 10, g] = b1
    return c + A @ B


def _r_layout(X, lane):
    g, t = lane >> 2, lane & 3

    def row(r):
        return np.concatenate([X[r, 8 * t:8 * t + 8], X[r, 32 + 8 * t:32 + 8 * t + 8]]).reshape(8, 2)
    return row(g), row(g + 8)


def test_r_layout_gives_x_times_y_transposed():
    rng = np.random.default_rng(0)
    Q, K = rng.standard_normal((16, 64)), rng.standard_normal((16, 64))
    S = np.zeros((16, 16))
    for nt in range(2):
        c = np.zeros((16, 8))
        for kk in range(4):
            a, b = {}, {}
            for lane in range(32):
                qlo, qhi = _r_layout(Q, lane)
                klo, khi = _r_layout(K, lane)
                y = klo if nt == 0 else khi
                a[lane] = (qlo[2 * kk], qhi[2 * kk], qlo[2 * kk + 1], qhi[2 * kk + 1])
                b[lane] = (y[2 * kk], y[2 * kk + 1])
            c = _mma_m16n8k16(c, a, b)
        S[:, 8 * nt:8 * nt + 8] = c
    assert np.abs(S - Q @ K.T).max() < 1e-12


def test_p_layout_and_output_column_permutation():
    rng = np.random.default_rng(1)
    P, V = rng.standard_normal((16, 16)), rng.standard_normal((16, 64))
    O = np.zeros((16, 64))
    for nt in range(8):
        a, b = {}, {}
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            # the accumulator layout of a 16x16 product IS this A-fragment layout (how P feeds P.V without a shuffle)
            a[lane] = (P[g, 2 * t:2 * t + 2], P[g + 8, 2 * t:2 * t + 2], P[g, 2 * t + 8:2 * t + 10], P[g + 8, 2 * t + 8:2 * t + 10])
            w = [V[r, 8 * g:8 * g + 8] for r in (2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9)]
            b[lane] = (np.array([w[0][nt], w[1][nt]]), np.array([w[2][nt], w[3][nt]]))
        c = _mma_m16n8k16(np.zeros((16, 8)), a, b)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            O[g, 16 * t + nt], O[g, 16 * t + 8 + nt] = c[g, 2 * t], c[g, 2 * t + 1]
            O[g + 8, 16 * t + nt], O[g + 8, 16 * t + 8 + nt] = c[g + 8, 2 * t], c[g + 8, 2 * t + 1]
    assert np.abs(O - P @ V).max() < 1e-12
