"""CPU model of the small-head-dim attention kernel's algorithm (csrc/attention_small.cu): offset folded into the S accumulator
through an augmented K step, lazily updated stale maximum with a carry for the S tile that was already issued, overflow
detection on the packed bf16 P' words (head_dim 16) or the tile's row sum (head_dim 32), row sums taken from the bf16-rounded
P' (ones row in V^T, head_dim 16).  The model follows the kernel's control flow tile by tile (64 keys, warps of 32 rows, S
issued two tiles ahead) in plain torch and must reproduce softmax(QK^T/sqrt(hd))V for inputs whose row maximum keeps growing,
shrinks, or is several hundred.  It is test code: the product path never runs it."""
import math

import pytest
import torch

K_SHIFT, K_LAZY, BKV = 8.0, 6.0, 64


def _bf(x):
    return x.to(torch.bfloat16).float()


def model_attention_small(q, k, v, lt):
    """q, k, v [N, hd] fp32 (bf16-representable).  Returns (out [N, hd], lse_log2 [N], number of slow tiles, tiles)."""
    N, hd = q.shape
    sl2 = math.log2(math.e) / math.sqrt(hd)
    mant, ex = math.frexp(sl2)
    c, qscale = mant * 2.0, 2.0 ** (ex - 1)
    qs = q * qscale
    assert torch.equal(_bf(qs), qs)           # the in-kernel Q scaling is exact (power of two)
    shift = K_SHIFT / c
    nkv = (N + BKV - 1) // BKV
    out, lse = torch.zeros(N, hd), torch.zeros(N)
    n_slow = n_tiles = 0
    for w0 in range(0, N, 32):                # one warp = 32 query rows; path decisions are warp-wide
        rows = slice(w0, min(w0 + 32, N))
        R = rows.stop - rows.start
        E = _bf(torch.full((R,), shift))
        E_hist = [E.clone()]                  # E_hist[i] = offset after tile i-1 has been processed
        O = torch.zeros(R, hd, dtype=torch.float64)
        l = torch.zeros(R, dtype=torch.float64)
        carry = torch.zeros(R)
        for j in range(nkv):
            ks = slice(j * BKV, min((j + 1) * BKV, N))
            e_used = E_hist[max(j - 1, 0)]    # S(q, j) is issued once tile j-2 has been processed (tiles 0, 1: prologue)
            S = qs[rows] @ k[ks].T - e_used[:, None]
            assert torch.equal(carry, E - e_used), "carry must equal the offset drift of the already issued S tile"
            tail = (j + 1) * BKV > N
            slow = tail or j == 0 or bool((carry != 0).any())
            if not slow:
                p32 = torch.exp2(c * S)
                P = _bf(p32)
                if lt:
                    slow = bool(((P >= 2.0) | (P < 0) | ~torch.isfinite(P)).any())
                else:
                    slow = not bool((p32.sum(1) < 32.0).all())
            if slow:
                tmax = S.max(1).values - carry
                upd = torch.ones(R, dtype=torch.bool) if j == 0 else tmax > -shift + K_LAZY / c
                e_new = torch.where(upd, _bf(E + tmax + shift), E)
                delta = e_new - E
                alpha = torch.exp2(-(delta * c)).double()
                O *= alpha[:, None]
                l *= alpha
                p32 = torch.exp2(c * S - ((carry + delta) * c)[:, None])
                P = _bf(p32)
                E, carry = e_new, delta
                n_slow += 1
            else:
                carry = torch.zeros(R)
            n_tiles += 1
            O += P.double() @ v[ks].double()
            l += (P if lt else p32).double().sum(1)
            E_hist.append(E.clone())
        out[rows] = (O / l[:, None]).float()
        lse[rows] = c * E + torch.log2(l).float()
    return out, lse, n_slow, n_tiles


@pytest.mark.parametrize("hd,N,mode", [(16, 512, "plain"), (16, 520, "grow"), (32, 392, "grow"), (16, 384, "shrink"), (32, 256, "big"),
                                       (16, 136, "grow"), (32, 1024, "plain")])
def test_model_matches_softmax(hd, N, mode):
    g = torch.Generator().manual_seed(hd * 1000 + N)
    q, k, v = (torch.randn(N, hd, generator=g) for _ in range(3))
    if mode == "grow":
        k = k * torch.linspace(0.25, 14.0, N)[:, None]
    elif mode == "shrink":
        k = k * torch.linspace(14.0, 0.05, N)[:, None]
    elif mode == "big":
        q, k = q * 6.0, k * 6.0
    q, k, v = _bf(q), _bf(k), _bf(v)
    out, lse, n_slow, n_tiles = model_attention_small(q, k, v, lt=(hd == 16))
    s = (q.double() @ k.double().T) / math.sqrt(hd)
    ref = (torch.softmax(s, dim=-1) @ v.double()).float()
    err = float((out - ref).norm() / ref.norm())
    assert torch.isfinite(out).all() and err < 1e-2, err
    lse_ref = (torch.logsumexp(s, dim=-1) * math.log2(math.e)).float()
    assert float((lse - lse_ref).abs().max()) < 2e-2 + 1e-5 * float(lse_ref.abs().max())
    if mode == "plain":   # unit-variance logits: after the two start-up tiles the slow path must be the exception
        assert n_slow <= 2 * (N // 32) + 0.1 * n_tiles, (n_slow, n_tiles)
