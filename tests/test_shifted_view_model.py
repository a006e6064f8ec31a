"""CPU model of conv_tcgen05.cu's data movement (flat padded index, shared zero row/column, shifted operand views,
blocked weight image order, epilogue decode), checked against torch's conv2d.  Pins the ALGORITHM of the tensor-core
engine where no GPU is available; the PTX-level behaviour is pinned by the GPU tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

KCK, MCTA = 32, 256


def pack_weights(w, NT):
    """[n_tile][pass][tap][kchunk][n][8]  (conv_tcgen05_pack_weights)."""
    cout, cin, k, _ = w.shape
    taps = k * k
    wf = w.reshape(cout, cin, taps)
    out = np.zeros((cout // NT, cin // KCK, taps, 4, NT, 8), np.float32)
    for nt in range(cout // NT):
        for c in range(cin // KCK):
            for t in range(taps):
                for kc in range(4):
                    out[nt, c, t, kc] = wf[nt * NT:(nt + 1) * NT, c * KCK + kc * 8:c * KCK + kc * 8 + 8, t]
    return out


def shifted_view_conv(x_nhwc, w, ksize):
    B, H, W, Cin = x_nhwc.shape
    cout = w.shape[0]
    NT = 128 if cout % 128 == 0 else (64 if cout % 64 == 0 else 32)
    pad = 1 if ksize == 3 else 0
    Wp, S, HW = W + pad, (H + pad) * (W + pad), H * W
    halo = Wp + 1 if ksize == 3 else 0
    P = MCTA + 2 * halo
    PA = P
    while PA % 8 != 2:
        PA += 1
    total = B * S
    wimg = pack_weights(w, NT)
    xf = x_nhwc.reshape(B * HW, Cin)
    out = np.zeros((B * HW, cout), np.float32)

    def decode(f):
        if f < 0 or f >= total:
            return -1, -1
        img, rem = divmod(f, S)
        row, col = divmod(rem, Wp)
        if row >= pad and col >= pad:
            return img, (row - pad) * (Wp - pad) + (col - pad)
        return img, -1

    for m0 in range(0, total, MCTA):
        for nt in range(cout // NT):
            acc = np.zeros((MCTA, NT), np.float32)
            for c in range(Cin // KCK):
                A = np.zeros((4, PA, 8), np.float32)             # [kchunk][pixel][8 ch]; pads stay zero
                for pixel in range(P):
                    img, pix = decode(m0 - halo + pixel)
                    if pix >= 0:
                        A[:, pixel, :] = xf[img * HW + pix, c * KCK:(c + 1) * KCK].reshape(4, 8)
                for t in range(ksize * ksize):
                    ky, kx = divmod(t, ksize)
                    delta = (ky - ksize // 2) * Wp + (kx - ksize // 2)
                    Bt = wimg[nt, c, t]                           # [4][NT][8]
                    for k16 in range(2):
                        for mt in range(2):
                            r0 = halo + mt * 128 + delta
                            a = A[2 * k16:2 * k16 + 2, r0:r0 + 128, :]        # the shifted view: 128 rows x 16 k
                            a = a.transpose(1, 0, 2).reshape(128, 16)
                            b = Bt[2 * k16:2 * k16 + 2].transpose(1, 0, 2).reshape(NT, 16)
                            acc[mt * 128:(mt + 1) * 128] += a @ b.T
            for r in range(MCTA):
                img, pix = decode(m0 + r)
                if pix >= 0:
                    out[img * HW + pix, nt * NT:(nt + 1) * NT] = acc[r]
    return out.reshape(B, H, W, cout)


@pytest.mark.parametrize("shape", [(3, 4, 4, 32, 32, 3), (2, 8, 8, 64, 64, 3), (1, 16, 16, 32, 128, 3), (5, 7, 7, 32, 32, 3),
                                   (2, 8, 8, 64, 96, 1), (3, 4, 4, 32, 384, 1)])
def test_shifted_view_formulation_equals_conv2d(shape):
    B, H, W, Cin, Cout, k = shape
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    ref = F.conv2d(x, w, padding=k // 2).permute(0, 2, 3, 1).numpy()
    got = shifted_view_conv(x.permute(0, 2, 3, 1).contiguous().numpy(), w.numpy(), k)
    assert np.abs(got - ref).max() < 1e-4


def test_garbage_row_fraction_is_as_documented():
    # MMA rows spent on pad positions = 1 - HW/S
    for hw, frac in ((32, 0.06), (16, 0.114), (8, 0.21), (4, 0.36)):
        assert abs((1 - hw * hw / (hw + 1) ** 2) - frac) < 0.005


# ---- the other geometries of conv_tcgen05.cu: GEO_DOWN (k4 s2 p1), GEO_UP (ConvTranspose k4 s2 p1), GEO_INIT (7x7 stem) ----
def _run_generic(B, S, Wv, halo_lo, halo_hi, n_pass, ntap, NT, n_tiles, fill_a, delta_of, wblock, store):
    """The CTA loop of the kernel with geometry-specific callbacks."""
    P = MCTA + halo_lo + halo_hi
    total = B * S
    for m0 in range(0, total, MCTA):
        for nt in range(n_tiles):
            acc = np.zeros((MCTA, NT), np.float32)
            for c in range(n_pass):
                A = np.zeros((4, P + 8, 8), np.float32)
                for pixel in range(P):
                    f = m0 - halo_lo + pixel
                    if 0 <= f < total:
                        img, rem = divmod(f, S)
                        row, col = divmod(rem, Wv)
                        v = fill_a(img, row, col, c)
                        if v is not None:
                            A[:, pixel, :] = v.reshape(4, 8)
                for t in range(ntap):
                    d = delta_of(t, nt)
                    Bt = wblock(nt, c, t)                    # [4][NT][8]
                    for k16 in range(2):
                        for mt in range(2):
                            r0 = halo_lo + mt * 128 + d
                            a = A[2 * k16:2 * k16 + 2, r0:r0 + 128, :].transpose(1, 0, 2).reshape(128, 16)
                            b = Bt[2 * k16:2 * k16 + 2].transpose(1, 0, 2).reshape(NT, 16)
                            acc[mt * 128:(mt + 1) * 128] += a @ b.T
            for r in range(MCTA):
                f = m0 + r
                if f < total:
                    img, rem = divmod(f, S)
                    row, col = divmod(rem, Wv)
                    store(img, row, col, nt, acc[r])


@pytest.mark.parametrize("B,H,C,Cout", [(3, 8, 32, 32), (2, 16, 64, 64), (5, 4, 32, 64)])
def test_down_as_space_to_depth_2x2(B, H, C, Cout):
    W = H
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, 4, 4, generator=g) / (C * 16) ** 0.5
    ref = F.conv2d(x, w, stride=2, padding=1).permute(0, 2, 3, 1).numpy()
    xn, wn = x.permute(0, 2, 3, 1).numpy(), w.numpy()
    NT = 64 if Cout % 64 == 0 else 32
    Wv = W // 2 + 1
    S = (H // 2 + 1) * Wv
    out = np.zeros((B, H // 2, W // 2, Cout), np.float32)

    def fill_a(img, u, v, c):
        vals = np.zeros(32, np.float32)
        for i in range(32):
            vc = c * 32 + i
            sub, ci = divmod(vc, C)
            iy, ix = 2 * u - 1 + (sub >> 1), 2 * v - 1 + (sub & 1)
            if 0 <= iy < H and 0 <= ix < W:
                vals[i] = xn[img, iy, ix, ci]
        return vals

    def wblock(nt, c, t):
        blk = np.zeros((4, NT, 8), np.float32)
        for kc in range(4):
            for e in range(8):
                vc = c * 32 + kc * 8 + e
                sub, ci = divmod(vc, C)
                ky, kx = 2 * (t >> 1) + (sub >> 1), 2 * (t & 1) + (sub & 1)
                blk[kc, :, e] = wn[nt * NT:(nt + 1) * NT, ci, ky, kx]
        return blk

    def store(img, u, v, nt, row):
        if u < H // 2 and v < W // 2:
            out[img, u, v, nt * NT:(nt + 1) * NT] = row

    _run_generic(B, S, Wv, 0, Wv + 1, 4 * C // 32, 4, NT, Cout // NT, fill_a, lambda t, nt: (t >> 1) * Wv + (t & 1), wblock, store)
    assert np.abs(out - ref).max() < 1e-4


@pytest.mark.parametrize("B,H,C,Cout", [(3, 4, 32, 32), (2, 8, 64, 32), (1, 16, 32, 64)])
def test_up_as_four_subpixel_phases(B, H, C, Cout):
    W = H
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, Cout, 4, 4, generator=g) / (C * 4) ** 0.5
    ref = F.conv_transpose2d(x, w, stride=2, padding=1).permute(0, 2, 3, 1).numpy()
    xn, wn = x.permute(0, 2, 3, 1).numpy(), w.numpy()
    NT = 64 if Cout % 64 == 0 else 32
    tpp = Cout // NT
    Wv = W + 1
    S = (H + 1) * Wv
    out = np.zeros((B, 2 * H, 2 * W, Cout), np.float32)

    def dydx(t, ph):
        py, px, a, b = ph >> 1, ph & 1, t >> 1, t & 1
        dy = (0 if a else 1) if py else (-1 if a else 0)
        dx = (0 if b else 1) if px else (-1 if b else 0)
        return dy, dx

    def fill_a(img, row, col, c):
        if row >= 1 and col >= 1:
            return xn[img, row - 1, col - 1, c * 32:(c + 1) * 32].copy()
        return None

    def wblock(nt, c, t):
        ph, ct = divmod(nt, tpp)
        dy, dx = dydx(t, ph)
        ky, kx = (ph >> 1) + 1 - 2 * dy, (ph & 1) + 1 - 2 * dx
        blk = np.zeros((4, NT, 8), np.float32)
        for kc in range(4):
            for e in range(8):
                blk[kc, :, e] = wn[c * 32 + kc * 8 + e, ct * NT:(ct + 1) * NT, ky, kx]
        return blk

    def delta(t, nt):
        dy, dx = dydx(t, nt // tpp)
        return dy * Wv + dx

    def store(img, row, col, nt, vals):
        ph, ct = divmod(nt, tpp)
        if row >= 1 and col >= 1:
            out[img, 2 * (row - 1) + (ph >> 1), 2 * (col - 1) + (ph & 1), ct * NT:(ct + 1) * NT] = vals

    _run_generic(B, S, Wv, Wv + 1, Wv + 1, C // 32, 4, NT, 4 * tpp, fill_a, delta, wblock, store)
    assert np.abs(out - ref).max() < 1e-4


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 8, 3, 32), (3, 16, 1, 32), (1, 32, 3, 64)])
def test_stem_7x7_with_kx_packed_into_channels(B, H, Cin, Cout):
    W = H
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 7, 7, generator=g) / (Cin * 49) ** 0.5
    ref = F.conv2d(x, w, padding=3).permute(0, 2, 3, 1).numpy()
    xn, wn = x.numpy(), w.numpy()
    NT = 64 if Cout % 64 == 0 else 32
    Wv, S = W, (H + 3) * W
    out = np.zeros((B, H, W, Cout), np.float32)

    def fill_a(img, row, col, c):
        if row < 3:
            return None
        vals = np.zeros(32, np.float32)
        for vc in range(32):
            kx, ch = divmod(vc, Cin)
            ix = col + kx - 3
            if kx < 7 and 0 <= ix < W:
                vals[vc] = xn[img, ch, row - 3, ix]
        return vals

    def wblock(nt, c, t):
        blk = np.zeros((4, NT, 8), np.float32)
        for kc in range(4):
            for e in range(8):
                kx, ch = divmod(kc * 8 + e, Cin)
                if kx < 7:
                    blk[kc, :, e] = wn[nt * NT:(nt + 1) * NT, ch, t, kx]
        return blk

    def store(img, row, col, nt, vals):
        if row >= 3:
            out[img, row - 3, col, nt * NT:(nt + 1) * NT] = vals

    _run_generic(B, S, Wv, 3 * Wv, 3 * Wv, 1, 7, NT, Cout // NT, fill_a, lambda t, nt: (t - 3) * Wv, wblock, store)
    assert np.abs(out - ref).max() < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# TMA operand feed (round 2): the im2col traversal and the SWIZZLE_64B tile, modelled on the CPU
# ---------------------------------------------------------------------------------------------------------------------
def im2col_window(x_nhwc, tn, th, tw, P, lower, stride=1, off=(0, 0)):
    """What ONE cp.async.bulk.tensor...im2col copy delivers (tools/micro/im2col_umma_test.cu pins this on the GPU): P base positions of
    the bounding box [lower, dim - 1] x [lower, dim - 1] (upper corner 0) in (w, h, n) order starting at (tw, th, tn), step `stride`;
    element = x[n][h + off_h][w + off_w], zero outside the tensor."""
    B, H, W, C = x_nhwc.shape
    out = np.zeros((P, C), np.float32)
    n, h, w = tn, th, tw
    for i in range(P):
        hh, ww = h + off[1], w + off[0]
        if 0 <= n < B and 0 <= hh < H and 0 <= ww < W:
            out[i] = x_nhwc[n, hh, ww]
        w += stride
        if w > W - 1:
            w, h = lower, h + stride
            if h > H - 1:
                h, n = lower, n + 1
    return out


def tile_coords(m0, halo_lo, S, Wv, pad, down):
    """conv_tcgen05.cu tma_tile_coords: tensor coordinates of the window's first flat position (image -1 = zeros in front)."""
    f = m0 - halo_lo
    tn = f // S if f >= 0 else -1
    rem = f - tn * S
    row, col = divmod(rem, Wv)
    return (tn, 2 * row - 1, 2 * col - 1) if down else (tn, row - pad, col - pad)


@pytest.mark.parametrize("H,B,m0", [(8, 3, 0), (8, 3, 128), (4, 9, 128), (16, 2, 256), (5, 4, 0)])
def test_im2col_window_is_the_flat_padded_window_3x3(H, B, m0):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, H, H, 4)).astype(np.float32)
    pad, Wv = 1, H + 1
    S, halo = (H + 1) * Wv, Wv + 1
    P = 128 + 2 * halo
    tn, th, tw = tile_coords(m0, halo, S, Wv, pad, False)
    got = im2col_window(x, tn, th, tw, P, lower=-1)
    for i in range(P):                      # flat padded index: row 0 / col 0 of every image block are the shared pads
        f = m0 - halo + i
        want = np.zeros(4, np.float32)
        if f >= 0:
            img, rem = divmod(f, S)
            row, col = divmod(rem, Wv)
            if img < B and row >= 1 and col >= 1:
                want = x[img, row - 1, col - 1]
        assert np.array_equal(got[i], want), (i, f)


@pytest.mark.parametrize("H,B,m0,sub", [(8, 3, 0, 0), (8, 3, 0, 3), (16, 2, 128, 1), (4, 20, 128, 2)])
def test_im2col_window_of_the_k4s2_form(H, B, m0, sub):
    """Virtual position (u, v) of the (H/2+1) x (W/2+1) grid holds pixel (2u-1+sy, 2v-1+sx): traversal stride 2 from lower corner -1,
    the sub-position (sy, sx) is the instruction's (w, h) offset."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, H, H, 4)).astype(np.float32)
    Wv = H // 2 + 1
    S, P = Wv * Wv, 96
    sy, sx = sub >> 1, sub & 1
    tn, th, tw = tile_coords(m0, 0, S, Wv, 0, True)
    got = im2col_window(x, tn, th, tw, P, lower=-1, stride=2, off=(sx, sy))
    for i in range(P):
        img, rem = divmod(m0 + i, S)
        u, v = divmod(rem, Wv)
        iy, ix = 2 * u - 1 + sy, 2 * v - 1 + sx
        want = x[img, iy, ix] if img < B and 0 <= iy < H and 0 <= ix < H else np.zeros(4, np.float32)
        assert np.array_equal(got[i], want), (i, img, u, v)


def test_swizzle64_chunk_of_a_producer_thread_is_constant():
    """SWIZZLE_64B tile [row][64 B]: 16-byte chunk position pc of row r holds k-chunk pc ^ ((r >> 1) & 3) (address bits 4-5 ^= bits 7-8 of a
    1 KB aligned buffer).  A producer thread owns chunk position pc of rows px0 + 64 j (8 warps) / + 96 j / + 128 j: one k-chunk for all."""
    for step in (64, 96, 128):
        for px0 in range(step):
            for pc in range(4):
                ks = {pc ^ (((px0 + step * j) >> 1) & 3) for j in range(8)}
                assert len(ks) == 1
    # and the address form: byte offset r * 64 + pc * 16, XOR of bits [4:6) with bits [7:9)
    for r in range(64):
        for kc in range(4):
            addr = r * 64 + kc * 16
            sw = addr ^ (((addr >> 7) & 3) << 4)
            assert sw == r * 64 + ((kc ^ ((r >> 1) & 3)) * 16)
