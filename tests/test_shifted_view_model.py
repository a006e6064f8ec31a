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
