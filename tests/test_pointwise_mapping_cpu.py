"""Index algebra of the mma.sync pointwise kernels (csrc/conv_pointwise.cu: pw_small_oc_mma_kernel / pw_small_ic_mma_kernel),
emulated on the CPU with the PTX fragment layouts of mma.m16n8k16 / m16n8k8 (lane = 4 g + t):
  A (16 x 16): a0 = (row g, k 2t..2t+1), a1 = (row g+8, same k), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
  B (16 x 8):  b0 = (k 2t..2t+1, col g), b1 = (k 2t+8.., col g)        C (16 x 8): c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = row g+8
The kernels feed each lane's 16-byte chunk of 8 consecutive channels to two k-steps (the k index of a dot product is
arbitrary as long as both operands agree) and permute the output-channel tiles so that a lane owns 8 consecutive channels."""
import numpy as np


def mma16816(afr, bfr):
    a, b = np.zeros((16, 16)), np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            a[g, 2 * t + e], a[g + 8, 2 * t + e] = afr[lane][0][e], afr[lane][1][e]
            a[g, 2 * t + 8 + e], a[g + 8, 2 * t + 8 + e] = afr[lane][2][e], afr[lane][3][e]
            b[2 * t + e, g], b[2 * t + 8 + e, g] = bfr[lane][0][e], bfr[lane][1][e]
    c = a @ b
    return np.array([[c[l >> 2, 2 * (l & 3)], c[l >> 2, 2 * (l & 3) + 1], c[(l >> 2) + 8, 2 * (l & 3)], c[(l >> 2) + 8, 2 * (l & 3) + 1]]
                     for l in range(32)])


def mma1688(afr, bfr):
    a, b = np.zeros((16, 8)), np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            a[g, 2 * t + e], a[g + 8, 2 * t + e] = afr[lane][0][e], afr[lane][1][e]
            b[2 * t + e, g] = bfr[lane][e]
    c = a @ b
    return np.array([[c[l >> 2, 2 * (l & 3)], c[l >> 2, 2 * (l & 3) + 1], c[(l >> 2) + 8, 2 * (l & 3)], c[(l >> 2) + 8, 2 * (l & 3) + 1]]
                     for l in range(32)])


def test_small_oc_chunk_mapping():
    rng = np.random.default_rng(0)
    for ic, oc in [(32, 3), (64, 3), (128, 4), (64, 1)]:
        x, w = rng.standard_normal((16, ic)), rng.standard_normal((oc, ic))
        acc = np.zeros((32, 4))
        for q in range(ic // 32):
            for half in range(2):                                   # the chunk's elements [4 half, 4 half + 4) form k-step 2q + half
                afr, bfr = np.zeros((32, 4, 2)), np.zeros((32, 2, 2))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    ch = q * 32 + t * 8 + half * 4
                    afr[lane][0], afr[lane][2] = x[g, ch:ch + 2], x[g, ch + 2:ch + 4]
                    afr[lane][1], afr[lane][3] = x[g + 8, ch:ch + 2], x[g + 8, ch + 2:ch + 4]
                    if g < oc:
                        bfr[lane][0], bfr[lane][1] = w[g, ch:ch + 2], w[g, ch + 2:ch + 4]
                acc += mma16816(afr, bfr)
        y = np.zeros((16, oc))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e in range(2):
                if 2 * t + e < oc:
                    y[g, 2 * t + e], y[g + 8, 2 * t + e] = acc[lane][e], acc[lane][2 + e]
        assert np.allclose(y, x @ w.T), (ic, oc)


def test_small_ic_tile_permutation():
    rng = np.random.default_rng(1)
    for ic, oc in [(3, 32), (3, 64), (4, 128), (2, 64), (1, 32)]:
        x, w = rng.standard_normal((16, ic)), rng.standard_normal((oc, ic))
        y = np.zeros((16, oc))
        afr = np.zeros((32, 2, 2))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e in range(2):
                if 2 * t + e < ic:
                    afr[lane][0][e], afr[lane][1][e] = x[g, 2 * t + e], x[g + 8, 2 * t + e]
        for q in range(oc // 32):
            for j in range(4):
                bfr = np.zeros((32, 2))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    o = q * 32 + (g >> 1) * 8 + j * 2 + (g & 1)      # channel of column n = g of tile (q, j)
                    for e in range(2):
                        if 2 * t + e < ic:
                            bfr[lane][e] = w[o, 2 * t + e]
                c = mma1688(afr, bfr)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    base = q * 32 + t * 8 + j * 2                    # the lane's 8 consecutive channels: 32 q + 8 t + [0, 8)
                    y[g, base:base + 2], y[g + 8, base:base + 2] = c[lane][0:2], c[lane][2:4]
        assert np.allclose(y, x @ w.T), (ic, oc)
