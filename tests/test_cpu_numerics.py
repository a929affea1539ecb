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
