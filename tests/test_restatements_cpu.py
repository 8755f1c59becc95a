"""CPU checks of the algebraic restatements the CUDA kernels rely on (fp64 torch, no GPU, no library call).

Each kernel that does NOT evaluate the reference's formula literally has its identity pinned here against the plain
torch op the reference uses, so a parity failure on the GPU can be told apart from a wrong derivation:

  * narrow heads (conv_igemm_t_kernel<1> + head_stencil_kernel): Conv3d(k=3, pad=1) with tiny Cout ==
    depth taps accumulated in a GEMM over depth-shifted inputs + a 9-tap in-plane stencil of the product rows
  * ConvTranspose3d(k=(3,4,4), s=(1,2,2), p=1) (reference models/unet3d.py:218, models/vae.py:86) == four
    output-parity classes of 3x2x2-tap ordinary convolutions (conv_host.cu CONV_UPT tap table)
  * strided Conv3d(k=(3,4,4), s=(1,2,2), p=1) == 48 taps over four input-parity views (CONV_DOWN tap table)
  * TemporalAttention fold (gn_res_tsum -> attn_gemm -> add_bcast_t): sum_t GN(x) from raw depth sums, also when the
    depth axis is split into partial sums
  * the persistent kernels' unit -> (tile, k-range) enumeration (single CTA, CTA pairs, split-K) covers every
    (m-tile, class, n-tile, k-step) exactly once
"""
import itertools

import pytest
import torch
import torch.nn.functional as F

torch.manual_seed(0)


def test_head_conv_equals_depth_gemm_plus_inplane_stencil():
    N, Cin, Cout, D, H, W = 2, 6, 3, 4, 5, 7
    x = torch.randn(N, Cin, D, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, 3, dtype=torch.float64)
    b = torch.randn(Cout, dtype=torch.float64)
    ref = F.conv3d(x, w, b, padding=1)
    # P[(kh,kw,co)][n,d,h,w] = sum_kd sum_c w[co,c,kd,kh,kw] * x[n,c,d+kd-1,h,w]   (zero outside the volume)
    xp = F.pad(x, (0, 0, 0, 0, 1, 1))
    P = torch.zeros(3, 3, N, Cout, D, H, W, dtype=torch.float64)
    for kh, kw, kd in itertools.product(range(3), range(3), range(3)):
        P[kh, kw] += torch.einsum("oc,ncdhw->nodhw", w[:, :, kd, kh, kw], xp[:, :, kd:kd + D])
    # out[q] = bias + sum_{kh,kw} P[(kh,kw)][q + (kh-1, kw-1)] over the in-bounds taps
    out = b.view(1, -1, 1, 1, 1).expand(N, Cout, D, H, W).clone()
    Pp = F.pad(P, (1, 1, 1, 1))
    for kh, kw in itertools.product(range(3), range(3)):
        out += Pp[kh, kw][..., kh:kh + H, kw:kw + W]
    assert torch.allclose(out, ref, atol=1e-12)


def test_transposed_conv_equals_four_parity_classes():
    N, Cin, Cout, D, H, W = 1, 3, 2, 3, 4, 5
    x = torch.randn(N, Cin, D, H, W, dtype=torch.float64)
    w = torch.randn(Cin, Cout, 3, 4, 4, dtype=torch.float64)
    ref = F.conv_transpose3d(x, w, None, stride=(1, 2, 2), padding=1)
    assert ref.shape == (N, Cout, D, 2 * H, 2 * W)
    # out(od, 2j+ph, 2i+pw) = sum in(od+1-kd, j+dh, i+dw) * w[ci][co][kd][kh][kw]
    #   ph=0: kh in {1 (dh 0), 3 (dh -1)};  ph=1: kh in {0 (dh +1), 2 (dh 0)}        (same table along w)
    ks = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    for ph, pw in itertools.product(range(2), range(2)):
        acc = torch.zeros(N, Cout, D, H, W, dtype=torch.float64)
        for kd in range(3):
            for (kh, dh), (kw, dw) in itertools.product(ks[ph], ks[pw]):
                dd = 1 - kd
                src = xp[:, :, 1 + dd:1 + dd + D, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W]
                acc += torch.einsum("co,ncdhw->nodhw", w[:, :, kd, kh, kw], src)
        out[..., ph::2, pw::2] = acc
    assert torch.allclose(out, ref, atol=1e-12)


def test_strided_conv_equals_taps_over_parity_views():
    N, Cin, Cout, D, H, W = 1, 3, 2, 3, 6, 8
    x = torch.randn(N, Cin, D, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 4, 4, dtype=torch.float64)
    ref = F.conv3d(x, w, None, stride=(1, 2, 2), padding=1)
    # input index 2*o + k - 1  ->  parity view p, offset dlt:  k=0:(1,-1) 1:(0,0) 2:(1,0) 3:(0,+1)
    par, dlt = (1, 0, 1, 0), (-1, 0, 0, 1)
    gH, gW = H // 2, W // 2
    out = torch.zeros(N, Cout, D, gH, gW, dtype=torch.float64)
    for kd, kh, kw in itertools.product(range(3), range(4), range(4)):
        view = F.pad(x[..., par[kh]::2, par[kw]::2], (1, 1, 1, 1, 1, 1))
        src = view[:, :, kd:kd + D, 1 + dlt[kh]:1 + dlt[kh] + gH, 1 + dlt[kw]:1 + dlt[kw] + gW]
        out += torch.einsum("oc,ncdhw->nodhw", w[:, :, kd, kh, kw], src)
    assert torch.allclose(out, ref, atol=1e-12)


@pytest.mark.parametrize("splits", [1, 2, 4])
def test_attention_fold_from_raw_depth_sums(splits):
    """x + Wp.(Wv.sum_t GN(x) + T.bv) + bp, with sum_t GN(x) = gamma*rstd*(sum_t x - T*mean) + T*beta rebuilt from
    per-split raw depth sums and the (sum, sumsq) statistics -- the data flow of gn_res_tsum -> attn_gemm."""
    B, C, T, H, W, G = 2, 16, 8, 3, 2, 4
    x = torch.randn(B, C, T, H, W, dtype=torch.float64)
    gn = torch.nn.GroupNorm(G, C).double()
    with torch.no_grad():
        gn.weight.normal_()
        gn.bias.normal_()
    Wv, bv = torch.randn(C, C, dtype=torch.float64), torch.randn(C, dtype=torch.float64)
    Wp, bp = torch.randn(C, C, dtype=torch.float64), torch.randn(C, dtype=torch.float64)
    with torch.no_grad():
        v = torch.einsum("oc,bcthw->bothw", Wv, gn(x)) + bv.view(1, -1, 1, 1, 1)
        ref = x + torch.einsum("oc,bchw->bohw", Wp, v.sum(2)).unsqueeze(2) + bp.view(1, -1, 1, 1, 1)
    # kernels: raw partial depth sums + raw statistics
    bounds = [s * T // splits for s in range(splits + 1)]
    tsum = torch.stack([x[:, :, bounds[s]:bounds[s + 1]].sum(2) for s in range(splits)], 1)  # [B][splits][C][H][W]
    xg = x.reshape(B, G, -1)
    n = xg.shape[-1]
    mean = xg.sum(-1) / n
    var = (xg * xg).sum(-1) / n - mean * mean
    rstd = (var + gn.eps).rsqrt()
    cpg = C // G
    mean_c, rstd_c = mean.repeat_interleave(cpg, 1), rstd.repeat_interleave(cpg, 1)
    with torch.no_grad():
        scl = gn.weight.view(1, -1) * rstd_c
        shf = T * (gn.bias.view(1, -1) - mean_c * scl)
        s = scl.view(B, C, 1, 1) * tsum.sum(1) + shf.view(B, C, 1, 1)
        g = torch.einsum("oc,bchw->bohw", Wp @ Wv, s) + (T * (Wp @ bv) + bp).view(1, -1, 1, 1)
        out = x + g.unsqueeze(2)
    assert torch.allclose(out, ref, atol=1e-10)


def _enumerate_units(m_tiles, n_tiles, nclass, splitk, ksteps, grid, pair):
    """host model of the unit walk in conv_igemm_kernel (conv_igemm.cuh: unit0 / unit_step / unit_k0 / unit_k1)"""
    m_units = (m_tiles + 1) // 2 if pair else m_tiles
    total_units = m_units * nclass * n_tiles * splitk
    workers = grid // 2 if pair else grid
    seen = []
    for worker in range(workers):
        for rank in range(2 if pair else 1):
            for unit in range(worker, total_units, workers):
                tile = unit // splitk
                m = 2 * (tile % m_units) + rank if pair else tile % m_units
                rest = tile // m_units
                cls, n = rest % nclass, rest // nclass
                k0 = (unit % splitk) * ksteps // splitk
                k1 = (unit % splitk + 1) * ksteps // splitk
                if m >= m_tiles:
                    continue  # padding half of an odd pair: loads are zero-filled, the epilogue skips it
                seen += [(m, cls, n, k) for k in range(k0, k1)]
    return seen


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("m_tiles,n_tiles,nclass,splitk,ksteps,grid", [
    (81, 1, 1, 1, 27, 148), (40, 2, 1, 1, 54, 148), (20, 1, 4, 1, 12, 148), (14, 2, 1, 5, 216, 140),
    (7, 2, 1, 3, 54, 42), (1, 1, 1, 1, 2, 2), (3, 4, 1, 2, 33, 8)])
def test_persistent_unit_walk_covers_every_k_step_once(pair, m_tiles, n_tiles, nclass, splitk, ksteps, grid):
    seen = _enumerate_units(m_tiles, n_tiles, nclass, splitk, ksteps, grid, pair)
    want = set(itertools.product(range(m_tiles), range(nclass), range(n_tiles), range(ksteps)))
    assert len(seen) == len(want) and set(seen) == want
