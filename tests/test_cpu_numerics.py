"""The numerics argument behind the tensor-core paths, checked in numpy -- no GPU.

f32 operands are split into bf16 planes x = p0 + p1 + p2 (round to nearest even at every step, csrc/vb_tc.cuh:tc_split3), the
tensor cores multiply planes (exact products in an f32 accumulator).  Claims:
  1. the split is exact: p0 + p1 + p2 == x for every finite f32 in the range the model produces;
  2. weights are bf16, so activation x weight needs the three plane products and nothing else (GEMM);
  3. f32 x f32 (attention: q.k and p.v) needs the six plane products a0b0, a0b1, a1b0, a1b1, a0b2, a2b0; what is dropped is below
     2^-24 of |a||b|, i.e. below the rounding of the f32 product itself.
"""
import numpy as np


def bf16_rne(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split3(x):
    x = np.asarray(x, dtype=np.float32)
    p0 = bf16_rne(x)
    r = (x - p0).astype(np.float32)            # exact in f32: r and p0 share the leading bits
    p1 = bf16_rne(r)
    r2 = (r - p1).astype(np.float32)
    p2 = bf16_rne(r2)
    return p0, p1, p2


def sample(rng, n):
    return (rng.normal(size=n) * np.exp(rng.uniform(-12, 6, size=n))).astype(np.float32)


def test_three_planes_are_exact():
    rng = np.random.default_rng(0)
    x = np.concatenate([sample(rng, 200000), np.float32([0.0, 1.0, -1.0, 3.0e-5, 65504.0, 1.0 + 2.0 ** -23, 1.0 - 2.0 ** -24])])
    p0, p1, p2 = split3(x)
    assert np.array_equal((p0.astype(np.float64) + p1 + p2).astype(np.float32), x)
    assert np.array_equal(p0.astype(np.float64) + p1.astype(np.float64) + p2.astype(np.float64), x.astype(np.float64))
    # each plane is a bf16 value (low 16 bits clear), magnitudes fall by >= 2^8 per plane
    for p in (p0, p1, p2):
        assert not (p.view(np.uint32) & 0xFFFF).any()
    nz = p0 != 0
    assert (np.abs(p1[nz]) <= np.abs(p0[nz]) * 2.0 ** -8).all() and (np.abs(p2[nz]) <= np.abs(p0[nz]) * 2.0 ** -16).all()


def test_activation_times_bf16_weight_needs_three_products():
    rng = np.random.default_rng(1)
    x = sample(rng, 100000)
    w = bf16_rne(sample(rng, 100000))
    p = split3(x)
    exact = x.astype(np.float64) * w.astype(np.float64)
    planes = sum(pi.astype(np.float64) * w.astype(np.float64) for pi in p)
    assert np.array_equal(planes, exact)                                 # bf16 x bf16 products are exact, and they add up exactly
    two = sum(pi.astype(np.float64) * w.astype(np.float64) for pi in p[:2])
    rel = np.abs(two - exact) / np.maximum(np.abs(exact), 1e-300)
    assert rel.max() <= 2.0 ** -16 and rel.max() > 2.0 ** -19            # what the third plane removes ("2^-17 split error")


def test_six_products_for_f32_times_f32():
    rng = np.random.default_rng(2)
    a, b = sample(rng, 200000), sample(rng, 200000)
    pa, pb = split3(a), split3(b)
    exact = a.astype(np.float64) * b.astype(np.float64)
    six = sum(pa[i].astype(np.float64) * pb[j].astype(np.float64) for i, j in ((0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)))
    rel = np.abs(six - exact) / np.maximum(np.abs(exact), 1e-300)
    assert rel.max() <= 2.0 ** -24                                       # below half an ulp of the f32 product
    three = sum(pa[i].astype(np.float64) * pb[j].astype(np.float64) for i, j in ((0, 0), (0, 1), (1, 0)))
    rel3 = np.abs(three - exact) / np.maximum(np.abs(exact), 1e-300)
    assert rel3.max() > 2.0 ** -18                                       # three products would not be enough


# ------------------------------------------------------------------------------------------------------------------------------
# The block schedule of the tcgen05 attention (csrc/vb_attn_tc.cu), restated in numpy: 128-query tiles, 64-key blocks from
# k_lo = floor64(max(0, g_first - W + 1)) to k_hi = min(g_last + 1, seq_k), mask only on blocks that are not interior
# (k0 >= lo_max and k0 + 63 <= hi_min), running maximum, deferred O = O*alpha(j-1) + (P V)(j-1), six plane products for S and P V.
# Checked against a float64 softmax over the reference's key range [max(0, g-W+1), min(g, seq_k-1)] (voxtral_kernels.c:412-482).
def attention_block_schedule(Q, K, V, window, q_offset, scale):
    seq_q, seq_k = Q.shape[0], K.shape[0]
    out = np.zeros_like(Q)
    six = ((0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0))

    def planes_matmul(A, B):                         # sum of the six plane products, f32 accumulate of exact products
        pa, pb = split3(A), split3(B)
        acc = np.zeros((A.shape[0], B.shape[0]), np.float64)
        for i, j in six:
            acc += pa[i].astype(np.float64) @ pb[j].astype(np.float64).T
        return acc.astype(np.float32)

    for q0 in range(0, seq_q, 128):
        rows = min(128, seq_q - q0)
        g_first, g_last = q_offset + q0, q_offset + q0 + rows - 1
        k_lo = max(0, g_first - window + 1) if window > 0 else 0
        k_lo = (k_lo // 64) * 64
        k_hi = min(g_last + 1, seq_k)
        nb = (k_hi - k_lo + 63) // 64 if k_hi > k_lo else 0
        lo_max = max(0, g_last - window + 1) if window > 0 else 0
        hi_min = min(g_first, seq_k - 1)
        g = q_offset + q0 + np.arange(rows)
        lo = np.maximum(0, g - window + 1) if window > 0 else np.zeros(rows, int)
        hi = np.minimum(g, seq_k - 1)
        m = np.full(rows, -1e30, np.float32); l = np.zeros(rows, np.float32)
        o = np.zeros((rows, Q.shape[1]), np.float32)
        pend = None                                   # (alpha of block j-1, (P V)(j-1))
        for j in range(nb):
            k0 = k_lo + j * 64
            kb = np.zeros((64, K.shape[1]), np.float32); vb = np.zeros((64, V.shape[1]), np.float32)
            n = max(0, min(64, seq_k - k0)); kb[:n] = K[k0:k0 + n]; vb[:n] = V[k0:k0 + n]
            s = planes_matmul(Q[q0:q0 + rows], kb) * np.float32(scale)
            if not (k0 >= lo_max and k0 + 63 <= hi_min):
                c = k0 + np.arange(64)
                s = np.where((c[None, :] >= lo[:, None]) & (c[None, :] <= hi[:, None]), s, np.float32(-1e30))
            mn = np.maximum(m, s.max(axis=1))
            alpha = np.exp((m - mn).astype(np.float64)).astype(np.float32)
            p = np.where(s > -1e29, np.exp((s - mn[:, None]).astype(np.float64)), 0.0).astype(np.float32)
            l = l * alpha + p.sum(axis=1, dtype=np.float32)
            m = mn
            if pend is not None:
                o = o * pend[0][:, None] + pend[1]
            pend = (alpha, planes_matmul(p, vb.T.copy()))
        if pend is not None:
            o = o * pend[0][:, None] + pend[1]
        out[q0:q0 + rows] = o / np.where(l > 0, l, 1)[:, None]
    return out


def attention_reference(Q, K, V, window, q_offset, scale):
    out = np.zeros(Q.shape, np.float64)
    for i in range(Q.shape[0]):
        g = q_offset + i
        a = max(0, g - window + 1) if window > 0 else 0
        b = min(g, K.shape[0] - 1)
        s = (K[a:b + 1].astype(np.float64) @ Q[i].astype(np.float64)) * scale
        p = np.exp(s - s.max())
        out[i] = (p / p.sum()) @ V[a:b + 1].astype(np.float64)
    return out


def test_attention_block_schedule_matches_exact_softmax():
    rng = np.random.default_rng(5)
    for seq_q, seq_k, win, qoff in ((300, 1500, 750, 1200), (128, 128, 750, 0), (129, 1000, 750, 871), (33, 97, 40, 64),
                                    (200, 200, 100, 0), (70, 70, 750, 0), (150, 900, 750, 750), (260, 1010, 750, 750)):
        Q = rng.normal(size=(seq_q, 64)).astype(np.float32)          # one head
        K = rng.normal(size=(seq_k, 64)).astype(np.float32)
        V = rng.normal(size=(seq_k, 64)).astype(np.float32)
        a = attention_block_schedule(Q, K, V, win, qoff, 0.125)
        b = attention_reference(Q, K, V, win, qoff, 0.125)
        err = np.abs(a - b).max()
        assert err < 2e-6 * max(np.abs(b).max(), 1.0), (seq_q, seq_k, win, qoff, err)
