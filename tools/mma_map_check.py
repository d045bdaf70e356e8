"""CPU check of the index algebra of csrc/dense_stream.cu (no GPU needed).

Emulates mma.sync.m16n8k16 at the fragment level (PTX ISA layouts, g = lane >> 2, c = lane & 3) and replays the
per-thread addressing of the three kernels on small problems; compares with a direct product.  Values are small
integers so float arithmetic is exact.
"""
import numpy as np

rng = np.random.default_rng(0)


def mma(a, b, d):
    """a[lane][4][2], b[lane][2][2] (pairs = consecutive k), d[lane][4] accumulators."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, c = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * c + e] = a[lane][0][e]; A[g + 8, 2 * c + e] = a[lane][1][e]
            A[g, 2 * c + 8 + e] = a[lane][2][e]; A[g + 8, 2 * c + 8 + e] = a[lane][3][e]
            B[2 * c + e, g] = b[lane][0][e]; B[2 * c + 8 + e, g] = b[lane][1][e]
    C = A @ B
    for lane in range(32):
        g, c = lane >> 2, lane & 3
        d[lane][0] += C[g, 2 * c]; d[lane][1] += C[g, 2 * c + 1]
        d[lane][2] += C[g + 8, 2 * c]; d[lane][3] += C[g + 8, 2 * c + 1]


def pairs(v8):            # 16-byte piece of 8 elements -> 4 registers of 2
    return [v8[0:2], v8[2:4], v8[4:6], v8[6:8]]


def check_fwd(M=20, N=70, K=128):
    w = rng.integers(-3, 4, (N, K)).astype(float); x = rng.integers(-3, 4, (M, K)).astype(float)
    acc = np.zeros((M, N))
    xs = np.zeros((32, K)); xs[:M] = x
    for blk in range((N + 255) // 256):
        for warp in range(8):
            n_base = blk * 256 + warp * 32
            if n_base >= N + 32:
                continue
            d = [[[np.zeros(4) for _ in range(32)] for _ in range(4)] for _ in range(2)]   # [rg][bb][lane]
            for step in range(K // 32):
                for bb in range(4):
                    for rg in range(2):
                        for half in range(2):     # MMA #1 (regs x,y) / #2 (regs z,w)
                            a = []; b = []
                            for lane in range(32):
                                g, c = lane >> 2, lane & 3
                                rows = [min(n_base + rg * 16 + h * 8 + g, N - 1) for h in range(2)]
                                wv = [pairs(w[r, step * 32 + c * 8: step * 32 + c * 8 + 8]) for r in rows]
                                xv = pairs(xs[bb * 8 + g, step * 32 + c * 8: step * 32 + c * 8 + 8])
                                a.append([wv[0][2 * half], wv[1][2 * half], wv[0][2 * half + 1], wv[1][2 * half + 1]])
                                b.append([xv[2 * half], xv[2 * half + 1]])
                            mma(a, b, d[rg][bb])
            for rg in range(2):
                for bb in range(4):
                    for lane in range(32):
                        g, c = lane >> 2, lane & 3
                        for e in range(4):
                            n = n_base + rg * 16 + g + (e >> 1) * 8; bt = bb * 8 + 2 * c + (e & 1)
                            if n < N and bt < M:
                                acc[bt, n] += d[rg][bb][lane][e]
    assert np.array_equal(acc, x @ w.T), "fwd mapping wrong"


def check_dgrad(M=20, N=38, K=128):
    w = rng.integers(-3, 4, (N, K)).astype(float); dy = rng.integers(-3, 4, (M, N)).astype(float)
    acc = np.zeros((M, K))
    nsteps = (N + 15) // 16
    ys = np.zeros((32, nsteps * 16)); ys[:M, :N] = dy
    for warp in range(K // 64):
        col0 = warp * 64
        d = [[[np.zeros(4) for _ in range(32)] for _ in range(8)] for _ in range(2)]   # [t][j][lane]
        for step in range(nsteps):
            bp = []; bq = []
            for lane in range(32):
                g, c = lane >> 2, lane & 3
                r = [min(step * 16 + 2 * c + (j & 1) + (j >> 1) * 8, N - 1) for j in range(4)]
                raw = [w[rr, col0 + g * 8: col0 + g * 8 + 8] for rr in r]
                bp.append([[raw[0][j], raw[1][j]] for j in range(8)])      # prmt 0x5410 / 0x7632
                bq.append([[raw[2][j], raw[3][j]] for j in range(8)])
            for t in range(2):
                a = []
                for lane in range(32):
                    g, c = lane >> 2, lane & 3
                    k0 = step * 16 + 2 * c
                    a.append([ys[t * 16 + g, k0:k0 + 2], ys[t * 16 + g + 8, k0:k0 + 2],
                              ys[t * 16 + g, k0 + 8:k0 + 10], ys[t * 16 + g + 8, k0 + 8:k0 + 10]])
                for j in range(8):
                    mma(a, [[bp[l][j], bq[l][j]] for l in range(32)], d[t][j])
        for t in range(2):
            for lane in range(32):
                g, c = lane >> 2, lane & 3
                for hi in range(2):
                    b = t * 16 + g + hi * 8
                    if b >= M:
                        continue
                    for half in range(2):
                        e = hi * 2 + half
                        for j in range(8):
                            acc[b, col0 + 16 * c + half * 8 + j] += d[t][j][lane][e]
    assert np.array_equal(acc, dy @ w), "dgrad mapping wrong"


def check_k1(H=7, W=9, R=5, S=5, pt=2, pl=2, band=3):
    P, Q = H, W
    x = rng.integers(-2, 3, (H, W, 64)).astype(float); w = rng.integers(-2, 3, (R * S, 64)).astype(float)
    ref = np.zeros((P, Q))
    for p in range(P):
        for q in range(Q):
            for r in range(R):
                for s in range(S):
                    ih, iw = p - pt + r, q - pl + s
                    if 0 <= ih < H and 0 <= iw < W:
                        ref[p, q] += x[ih, iw] @ w[r * S + s]
    out = np.zeros((P, Q))
    pixp = ((band + R - 1) * W + 15) // 16 * 16 // 32 * 32 + 36
    for p0 in range(0, P, band):
        rows_out = min(band, P - p0); rows_in = rows_out + R - 1; npix = rows_in * W; h0 = p0 - pt
        t = np.full((R * S, pixp), np.nan)
        nblk = (npix + 15) // 16
        assert nblk * 16 <= pixp
        for mb in range(nblk):
            d = [[np.zeros(4) for _ in range(32)] for _ in range(4)]
            for nb in range(4):
                for h in range(2):
                    for m in range(2):
                        a = []; b = []
                        for lane in range(32):
                            g, c = lane >> 2, lane & 3
                            px = []
                            for hi in range(2):
                                i = mb * 16 + g + hi * 8; row = i // W; ih = h0 + row
                                ok = i < npix and 0 <= ih < H
                                px.append(pairs(x[ih, i - row * W, h * 32 + c * 8: h * 32 + c * 8 + 8]) if ok else pairs(np.zeros(8)))
                            tap = nb * 8 + g
                            wv = pairs(w[tap, h * 32 + c * 8: h * 32 + c * 8 + 8]) if tap < R * S else pairs(np.zeros(8))
                            a.append([px[0][2 * m], px[1][2 * m], px[0][2 * m + 1], px[1][2 * m + 1]])
                            b.append([wv[2 * m], wv[2 * m + 1]])
                        mma(a, b, d[nb])
            for nb in range(4):
                for lane in range(32):
                    g, c = lane >> 2, lane & 3
                    for e in range(4):
                        tap = nb * 8 + 2 * c + (e & 1); i = mb * 16 + g + (e >> 1) * 8
                        if tap < R * S:
                            t[tap, i] = d[nb][lane][e]
        for o in range(rows_out * Q):
            pr, q = o // Q, o % Q
            a = 0.0
            for r in range(R):
                base = (pr + r) * W + (q - pl)
                for s in range(S):
                    if 0 <= q - pl + s < W:
                        a += t[r * S + s, base + s]
            out[p0 + pr, q] = a
    assert np.array_equal(out, ref), "k1 mapping wrong"


if __name__ == "__main__":
    check_fwd(); check_fwd(M=32, N=300, K=64)
    check_dgrad(); check_dgrad(M=32, N=64, K=64)
    check_k1(); check_k1(H=5, W=6, R=3, S=3, pt=1, pl=1, band=8)
    print("mma index maps OK")
