"""CPU emulation of the warp-level index math used by the mma.sync kernels.

The attention core (sv_attention.cu) and the small-M linear (sv_gemm_rowgroup.cu) feed
m16n8k16 tensor-core fragments straight from 128-bit global loads, relying on (a) a
permutation of the reduction index applied to both operands and (b) a key permutation that
makes the S accumulator columns line up with a transposed-V load.  This test replays exactly
those index expressions with a numpy model of the PTX fragment layout and checks the result
against plain matmul/softmax, so a layout mistake is caught without a GPU.
"""
import numpy as np


def mma_16816(c, a_regs, b_regs):
    """c: [32 lanes][4] fp32; a_regs: [32][4][2] (4 regs of 2 bf16); b_regs: [32][2][2]."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = a_regs[lane][0]
        A[g + 8, 2 * t:2 * t + 2] = a_regs[lane][1]
        A[g, 2 * t + 8:2 * t + 10] = a_regs[lane][2]
        A[g + 8, 2 * t + 8:2 * t + 10] = a_regs[lane][3]
        B[2 * t:2 * t + 2, g] = b_regs[lane][0]
        B[2 * t + 8:2 * t + 10, g] = b_regs[lane][1]
    Dm = A @ B
    out = c.copy()
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane][0] += Dm[g, 2 * t]; out[lane][1] += Dm[g, 2 * t + 1]
        out[lane][2] += Dm[g + 8, 2 * t]; out[lane][3] += Dm[g + 8, 2 * t + 1]
    return out


def words(vec8):
    """A 16-byte load of 8 bf16 -> 4 words of 2 elements."""
    return [vec8[0:2], vec8[2:4], vec8[4:6], vec8[6:8]]


def test_rowgroup_linear_layout():
    rng = np.random.default_rng(0)
    K, M = 96, 5
    W = rng.standard_normal((16, K)); X = rng.standard_normal((M, K))
    c = np.zeros((32, 4))
    for ch in range(K // 32):
        a_regs1, b_regs1, a_regs2, b_regs2 = [], [], [], []
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            a = words(W[g, ch * 32 + 8 * t: ch * 32 + 8 * t + 8])
            b = words(W[g + 8, ch * 32 + 8 * t: ch * 32 + 8 * t + 8])
            xv = words(X[g, ch * 32 + 8 * t: ch * 32 + 8 * t + 8]) if g < M else [np.zeros(2)] * 4
            a_regs1.append([a[0], b[0], a[1], b[1]]); b_regs1.append([xv[0], xv[1]])
            a_regs2.append([a[2], b[2], a[3], b[3]]); b_regs2.append([xv[2], xv[3]])
        c = mma_16816(c, a_regs1, b_regs1)
        c = mma_16816(c, a_regs2, b_regs2)
    Y = np.zeros((8, 16))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        Y[2 * t, g] = c[lane][0]; Y[2 * t + 1, g] = c[lane][1]
        Y[2 * t, g + 8] = c[lane][2]; Y[2 * t + 1, g + 8] = c[lane][3]
    np.testing.assert_allclose(Y[:M], X @ W.T, rtol=1e-10, atol=1e-10)


def _attn_core_emulated(Q, Kmat, V, key_end, scale):
    """Q [16,D], Kmat [T,D], V [T,D] -> O [16,D] via the kernel's fragment walk (single block loop)."""
    D = Q.shape[1]
    Vt = np.zeros((D, ((key_end + 31) // 32) * 32)); Vt[:, :V.shape[0]] = V.T
    qa = [[None] * 32 for _ in range(D // 16)]
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for jj in range(D // 32):
            a = words(Q[g, 32 * jj + 8 * t: 32 * jj + 8 * t + 8]); b = words(Q[g + 8, 32 * jj + 8 * t: 32 * jj + 8 * t + 8])
            qa[2 * jj][lane] = [a[0], b[0], a[1], b[1]]
            qa[2 * jj + 1][lane] = [a[2], b[2], a[3], b[3]]
    acc = [np.zeros((32, 4)) for _ in range(D // 8)]
    m = np.full((32, 2), -np.inf); l = np.zeros((32, 2))
    for kb in range(0, key_end, 32):
        s = []
        for j in range(4):
            sj = np.zeros((32, 4))
            for jj in range(D // 32):
                b1, b2 = [], []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    key = min(kb + 8 * (g >> 1) + 2 * j + (g & 1), key_end - 1)
                    w = words(Kmat[key, 32 * jj + 8 * t: 32 * jj + 8 * t + 8])
                    b1.append([w[0], w[1]]); b2.append([w[2], w[3]])
                sj = mma_16816(sj, qa[2 * jj], b1)
                sj = mma_16816(sj, qa[2 * jj + 1], b2)
            s.append(sj)
        s = np.stack(s, axis=1)  # [32][4 tiles][4]
        for lane in range(32):
            t = lane & 3
            for j in range(4):
                for e in range(2):
                    valid = (kb + 8 * t + 2 * j + e) < key_end
                    s[lane, j, e] = s[lane, j, e] * scale if valid else -np.inf
                    s[lane, j, 2 + e] = s[lane, j, 2 + e] * scale if valid else -np.inf
        mx = np.stack([s[:, :, 0:2].reshape(32, -1).max(1), s[:, :, 2:4].reshape(32, -1).max(1)], 1)
        mx = mx.reshape(8, 4, 2).max(1, keepdims=True).repeat(4, 1).reshape(32, 2)   # quad_max
        mn = np.maximum(m, mx)
        corr = np.exp2(m - mn); m = mn
        p = s.copy()
        p[:, :, 0:2] = np.exp2(s[:, :, 0:2] - mn[:, None, 0:1]); p[:, :, 2:4] = np.exp2(s[:, :, 2:4] - mn[:, None, 1:2])
        l[:, 0] = l[:, 0] * corr[:, 0] + p[:, :, 0:2].reshape(32, -1).sum(1)
        l[:, 1] = l[:, 1] * corr[:, 1] + p[:, :, 2:4].reshape(32, -1).sum(1)
        pa = [[[p[lane, 2 * h, 0:2], p[lane, 2 * h, 2:4], p[lane, 2 * h + 1, 0:2], p[lane, 2 * h + 1, 2:4]]
               for lane in range(32)] for h in range(2)]
        for nd in range(D // 8):
            acc[nd][:, 0:2] *= corr[:, 0:1]; acc[nd][:, 2:4] *= corr[:, 1:2]
            b1, b2 = [], []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                w = words(Vt[8 * nd + g, kb + 8 * t: kb + 8 * t + 8])
                b1.append([w[0], w[1]]); b2.append([w[2], w[3]])
            acc[nd] = mma_16816(acc[nd], pa[0], b1)
            acc[nd] = mma_16816(acc[nd], pa[1], b2)
    lq = l.reshape(8, 4, 2).sum(1, keepdims=True).repeat(4, 1).reshape(32, 2)       # quad_sum
    O = np.zeros((16, D))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for nd in range(D // 8):
            O[g, 8 * nd + 2 * t: 8 * nd + 2 * t + 2] = acc[nd][lane, 0:2] / lq[lane, 0]
            O[g + 8, 8 * nd + 2 * t: 8 * nd + 2 * t + 2] = acc[nd][lane, 2:4] / lq[lane, 1]
    return O


def test_attention_core_layout():
    rng = np.random.default_rng(1)
    for D, T in ((64, 41), (128, 70)):
        Q = rng.standard_normal((16, D)); Kmat = rng.standard_normal((T, D)); V = rng.standard_normal((T, D))
        scale = 1.0 / np.sqrt(D)
        O = _attn_core_emulated(Q, Kmat, V, T, scale * np.log2(np.e))
        S = Q @ Kmat.T * scale
        P = np.exp(S - S.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
        np.testing.assert_allclose(O, P @ V, rtol=1e-9, atol=1e-9)
