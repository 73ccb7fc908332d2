#!/usr/bin/env python
"""CPU emulation of the index maps of csrc/conv_c3_mma.cu (fragment layouts of mma.m16n8k16 / ldmatrix, shared-memory
addressing, epilogue scatter) against the direct sums, in float64.  A development aid: it mirrors the kernel code
line by line so that a mapping mistake shows up here instead of on the GPU box.  Not part of the product or tests."""
import numpy as np

LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def mma(acc, a, b0, b1):
    """acc[32][4] += A(16x16) @ B(16x8) with per-lane fragments a[32][4][2], b0/b1[32][2]."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        A[g, 2 * t:2 * t + 2] = a[l][0]; A[g + 8, 2 * t:2 * t + 2] = a[l][1]
        A[g, 2 * t + 8:2 * t + 10] = a[l][2]; A[g + 8, 2 * t + 8:2 * t + 10] = a[l][3]
        B[2 * t:2 * t + 2, g] = b0[l]; B[2 * t + 8:2 * t + 10, g] = b1[l]
    C = A @ B
    for l in range(32):
        g, t = l >> 2, l & 3
        acc[l] += [C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1]]


def ldsm(mem, addrs, trans):
    """mem: flat array of elements; addrs[32]: element offset of the 8-element row each lane supplies. -> r[32][4][2]"""
    mats = [np.stack([mem[addrs[8 * j + r]:addrs[8 * j + r] + 8] for r in range(8)]) for j in range(4)]
    out = np.zeros((32, 4, 2))
    for l in range(32):
        for j in range(4):
            if trans:
                out[l, j] = [mats[j][2 * (l & 3), l >> 2], mats[j][2 * (l & 3) + 1, l >> 2]]
            else:
                out[l, j] = mats[j][l >> 2, 2 * (l & 3):2 * (l & 3) + 2]
    return out


def ref_down(x, w):
    N, H, W, _ = x.shape; K = w.shape[-1]
    xp = np.zeros((N, H + 3, W + 3, 3)); xp[:, 1:H + 1, 1:W + 1] = x
    y = np.zeros((N, H // 2, W // 2, K))
    for r in range(5):
        for s in range(5):
            y += np.einsum("npqc,ck->npqk", xp[:, r:r + H:2, s:s + W:2], w[r, s])
    return y


def emu_down(x, w):
    N, H, W, _ = x.shape; K = w.shape[-1]; Ho, Wo = H // 2, W // 2
    TH, TW, PR, PC, PROW, WROW = 8, 16, 19, 35, 106, 88
    y = np.full((N, Ho, Wo, K), np.nan)
    wf = w.reshape(75, K)
    for kb in range(0, K, 64):
        swt = np.zeros(64 * WROW)
        for e in range(75 * 64):
            tc, ch = e >> 6, e & 63
            swt[ch * WROW + (tc // 15) * 16 + tc % 15] = wf[tc, kb + ch]
        for n in range(N):
            for p0 in range(0, Ho, TH):
                for q0 in range(0, Wo, TW):
                    sp = np.zeros(PR * PROW)
                    i0, j0 = 2 * p0 - 1, 2 * q0 - 1
                    for e in range(PR * PC * 3):
                        a, xx = divmod(e, PC * 3)
                        i, j = i0 + a, j0 + xx // 3
                        sp[a * PROW + xx] = x[n, i, j, xx % 3] if (0 <= i < H and 0 <= j < W) else 0.0
                    for warp in range(8):
                        acc = np.zeros((8, 32, 4))
                        for r in range(5):
                            row = (2 * warp + r) * PROW
                            a = np.zeros((32, 4, 2))
                            for l in range(32):
                                g, t = l >> 2, l & 3
                                for i, wo in enumerate([3 * g + t, 3 * (g + 8) + t, 3 * g + t + 4, 3 * (g + 8) + t + 4]):
                                    a[l, i] = sp[row + 2 * wo: row + 2 * wo + 2]
                            for j in range(8):
                                b0 = np.zeros((32, 2)); b1 = np.zeros((32, 2))
                                for l in range(32):
                                    g, t = l >> 2, l & 3
                                    ch = (g >> 1) * 16 + 2 * j + (g & 1)
                                    o = ch * WROW + r * 16 + 2 * t
                                    b0[l] = swt[o:o + 2]; b1[l] = swt[o + 8:o + 10]
                                mma(acc[j], a, b0, b1)
                        p = p0 + warp
                        if p >= Ho:
                            continue
                        for l in range(32):
                            g, t = l >> 2, l & 3
                            for half in range(2):
                                q = q0 + g + 8 * half
                                if q >= Wo:
                                    continue
                                for j in range(8):
                                    y[n, p, q, kb + t * 16 + 2 * j] = acc[j][l][2 * half]
                                    y[n, p, q, kb + t * 16 + 2 * j + 1] = acc[j][l][2 * half + 1]
    return y


def ref_up(xs, w):
    """large[n,i,j,c] = sum small[n,p,q,k] w[r,s,c,k], i = 2p + r - 1."""
    N, Ho, Wo, K = xs.shape; H, W = 2 * Ho, 2 * Wo
    out = np.zeros((N, H + 3, W + 3, 3))
    for r in range(5):
        for s in range(5):
            out[:, r:r + H:2, s:s + W:2] += np.einsum("npqk,ck->npqc", xs, w[r, s])
    return out[:, 1:H + 1, 1:W + 1]


def emu_up(xs, w):
    N, Ho, Wo, K = xs.shape; H, W = 2 * Ho, 2 * Wo
    TH, TW, PRr, PCc, PIX, WROW = 8, 16, 10, 18, 72, 584
    out = np.full((N, H, W, 3), np.nan)
    for n in range(N):
        for p0 in range(0, Ho, TH):
            for q0 in range(0, Wo, TW):
                accs = np.zeros((8, 2, 32, 4))
                for kb in range(0, K, 64):
                    swu = np.zeros(16 * WROW)
                    for e in range(16 * 9 * 16):
                        k4, nb, nn = (e & 15) * 4, (e >> 4) % 9, e // (16 * 9)
                        cls, c = nn // 3, nn % 3
                        a, b = cls >> 1, cls & 1
                        r, s = a + 1 - 2 * (nb // 3 - 1), b + 1 - 2 * (nb % 3 - 1)
                        if nn < 12 and 0 <= r < 5 and 0 <= s < 5:
                            swu[nn * WROW + nb * 64 + k4: nn * WROW + nb * 64 + k4 + 4] = w[r, s, c, kb + k4: kb + k4 + 4]
                    sx = np.zeros(PRr * PCc * PIX)
                    for pix in range(PRr * PCc):
                        p, q = p0 - 1 + pix // PCc, q0 - 1 + pix % PCc
                        if 0 <= p < Ho and 0 <= q < Wo:
                            sx[pix * PIX: pix * PIX + 64] = xs[n, p, q, kb:kb + 64]
                    for warp in range(8):
                        for nb in range(9):
                            for kc in range(4):
                                addrs = []
                                for lane in range(32):
                                    lrow, lk = (lane & 7) + 8 * ((lane >> 3) & 1), 8 * (lane >> 4)
                                    py, px = warp + nb // 3, lrow + nb % 3
                                    addrs.append((py * PCc + px) * PIX + lk + kc * 16)
                                a = ldsm(sx, addrs, False)
                                for j in range(2):
                                    b0 = np.zeros((32, 2)); b1 = np.zeros((32, 2))
                                    for l in range(32):
                                        g, t = l >> 2, l & 3
                                        o = (g + 8 * j) * WROW + nb * 64 + kc * 16 + 2 * t
                                        b0[l] = swu[o:o + 2]; b1[l] = swu[o + 8:o + 10]
                                    mma(accs[warp, j], a, b0, b1)
                for warp in range(8):
                    mi = p0 + warp
                    if mi >= Ho:
                        continue
                    for l in range(32):
                        g, t = l >> 2, l & 3
                        for half in range(2):
                            mj = q0 + g + 8 * half
                            if mj >= Wo:
                                continue
                            for j in range(2):
                                nn = 8 * j + 2 * t
                                if nn >= 12:
                                    continue
                                arow, off = nn // 6, nn % 6
                                flat = out[n].reshape(-1)
                                base = ((2 * mi + arow) * W + 2 * mj) * 3 + off
                                flat[base] = accs[warp, j][l][2 * half]
                                flat[base + 1] = accs[warp, j][l][2 * half + 1]
    return out


def ref_wgrad(x, ys):
    N, H, W, _ = x.shape; K = ys.shape[-1]
    xp = np.zeros((N, H + 3, W + 3, 3)); xp[:, 1:H + 1, 1:W + 1] = x
    dw = np.zeros((5, 5, 3, K))
    for r in range(5):
        for s in range(5):
            dw[r, s] = np.einsum("npqc,npqk->ck", xp[:, r:r + H:2, s:s + W:2], ys)
    return dw


def emu_wgrad(x, ys, with_bias=False):
    """with_bias: the bias rider of c3m_wgrad_kernel -- the unused fourth channel slot of every staged IN-IMAGE pixel holds 1, and the
    discarded row (r = 1, s = 1, c = 3) of the GEMM then accumulates sum_pixels ys[pixel, k]; returns (dw, dbias)."""
    N, H, W, _ = x.shape; K = ys.shape[-1]; Ho, Wo = H // 2, W // 2
    TH, TW, PR, PC, YPIX = 8, 16, 19, 36, 72
    dw = np.zeros((75, K))
    dbias = np.zeros(K)
    for kb in range(0, K, 64):
        acc = np.zeros((8, 8, 32, 4))      # warp, jn, lane, 4
        for n in range(N):
            for p0 in range(0, Ho, TH):
                for q0 in range(0, Wo, TW):
                    i0, j0 = 2 * p0 - 1, 2 * q0 - 1
                    sp = np.zeros(PR * PC * 4)
                    for e in range(PR * PC):
                        a, b = divmod(e, PC)
                        i, j = i0 + a, j0 + b
                        if 0 <= i < H and 0 <= j < W:
                            sp[e * 4:e * 4 + 3] = x[n, i, j]
                            if with_bias:
                                sp[e * 4 + 3] = 1.0
                    sy = np.zeros(TH * TW * YPIX)
                    for pix in range(TH * TW):
                        p, q = p0 + pix // TW, q0 + pix % TW
                        if p < Ho and q < Wo:
                            sy[pix * YPIX: pix * YPIX + 64] = ys[n, p, q, kb:kb + 64]
                    for warp in range(8):
                        for ks in range(TH * TW // 16):
                            addrs = []
                            for lane in range(32):
                                l8, lm = lane & 7, lane >> 3
                                mblk = min(2 * warp + (lm & 1), 14)
                                a_r, a_s0 = mblk // 3, 2 * (mblk % 3)
                                px = 8 * (lm >> 1) + l8
                                addrs.append(((2 * ks + a_r) * PC + 2 * px + a_s0) * 4)
                            a = ldsm(sp, addrs, True)
                            for jn in range(0, 8, 2):
                                addrs = []
                                for lane in range(32):
                                    l8, lm = lane & 7, lane >> 3
                                    pix = ks * 16 + 8 * (lm & 1) + l8
                                    addrs.append(pix * YPIX + (jn + (lm >> 1)) * 8)
                                b = ldsm(sy, addrs, True)
                                mma(acc[warp, jn], a, b[:, 0], b[:, 1])
                                mma(acc[warp, jn + 1], a, b[:, 2], b[:, 3])
        for warp in range(8):
            for l in range(32):
                g, t = l >> 2, l & 3
                for half in range(2):
                    mb = 2 * warp + half
                    r, s, c = mb // 3, 2 * (mb % 3) + (g >> 2), g & 3
                    if with_bias and r == 1 and s == 1 and c == 3:
                        for jn in range(8):
                            dbias[kb + jn * 8 + 2 * t] += acc[warp, jn][l][2 * half]
                            dbias[kb + jn * 8 + 2 * t + 1] += acc[warp, jn][l][2 * half + 1]
                    if mb >= 15 or s >= 5 or c >= 3:
                        continue
                    row = (r * 5 + s) * 3 + c
                    for jn in range(8):
                        dw[row, kb + jn * 8 + 2 * t] += acc[warp, jn][l][2 * half]
                        dw[row, kb + jn * 8 + 2 * t + 1] += acc[warp, jn][l][2 * half + 1]
    return (dw.reshape(5, 5, 3, K), dbias) if with_bias else dw.reshape(5, 5, 3, K)


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    x = rs.randn(1, 20, 36, 3); w = rs.randn(5, 5, 3, 64)
    d = np.abs(emu_down(x, w) - ref_down(x, w)).max()
    print("down max err", d); assert d < 1e-9
    xs = rs.randn(1, 10, 18, 64)
    d = np.abs(emu_up(xs, w) - ref_up(xs, w)).max()
    print("up max err", d); assert d < 1e-9
    ys = rs.randn(1, 10, 18, 64)
    d = np.abs(emu_wgrad(x, ys) - ref_wgrad(x, ys)).max()
    print("wgrad max err", d); assert d < 1e-9
    dw, db = emu_wgrad(x, ys, with_bias=True)                  # the bias rider leaves dw untouched and yields sum_pixels ys
    d = max(np.abs(dw - ref_wgrad(x, ys)).max(), np.abs(db - ys.sum((0, 1, 2))).max())
    print("wgrad + bias rider max err", d); assert d < 1e-9
    print("ok")
