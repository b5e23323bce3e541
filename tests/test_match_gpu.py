"""GPU parity tests of the Hamming matchers vs the CPU oracle (bit-exact: distances and indices)."""
import numpy as np
import pytest

from viorb_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from viorb_b200 import api
    api.lib()
    return api


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


def np_hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_descriptor_distance(api, ctx, oracle):
    m = api.ORBmatcher(ctx=ctx)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    b[:10] = a[:10]
    b[10] = ~a[10]
    d = m.DescriptorDistance(a, b)
    for i in range(500):
        assert d[i] == oracle.descriptor_distance(a[i], b[i]) == np_hamming(a[i], b[i])
    assert d[0] == 0 and d[10] == 256
    assert m.DescriptorDistance(a[3], b[77]) == np_hamming(a[3], b[77])


@pytest.mark.parametrize("Q,M", [(1000, 50000), (1, 1), (5, 255), (129, 256), (300, 100001), (7, 0), (1, 200003), (2, 70001),
                                 (3, 65536), (4, 99999), (8, 123457), (9, 40000)])
def test_top2_vs_oracle(api, ctx, oracle, Q, M):
    m = api.ORBmatcher(ctx=ctx)
    dmap = synth.descriptor_map(max(M, 1), seed=1234)[:M]
    q = synth.queries_from_map(dmap, Q, seed=5) if M > 0 else synth.descriptor_map(Q, seed=9)
    got = m.hamming_top2(q, dmap)
    ref = oracle.hamming_top2(q, dmap, nthreads=8)
    for f in ("d1", "i1", "d2", "i2"):
        assert (got[f] == ref[f]).all(), f


def test_top2_ties_and_base(api, ctx, oracle):
    """many exact ties at the minimum: the lowest index must win (strict < scan, ORBmatcher.cc:216-225)"""
    m = api.ORBmatcher(ctx=ctx)
    rng = np.random.default_rng(2)
    base = rng.integers(0, 256, (16, 32)).astype(np.uint8)
    dmap = base[rng.integers(0, 16, 20000)]          # only 16 distinct descriptors
    q = base[:8].copy()
    q[4:] ^= 1
    got = m.hamming_top2(q, dmap, index_base=1000)
    ref = oracle.hamming_top2(q, dmap, index_base=1000)
    for f in ("d1", "i1", "d2", "i2"):
        assert (got[f] == ref[f]).all(), f
    assert (got["d1"][:4] == 0).all() and (got["d2"][:4] == 0).all() and (got["i1"] < got["i2"]).all()


def test_top2_sharded_merge(api, ctx, oracle):
    """map sharded in 8 parts + merge == one scan over the whole map (the multi-GPU data path on one GPU)"""
    torch = pytest.importorskip("torch")
    m = api.ORBmatcher(ctx=ctx)
    Q, M, P = 200, 40000, 8
    dmap = synth.descriptor_map(M, seed=77)
    q = synth.queries_from_map(dmap, Q, seed=6)
    ref = oracle.hamming_top2(q, dmap, nthreads=8)
    dq = torch.from_numpy(q).cuda()
    dm = torch.from_numpy(dmap).cuda()
    parts = torch.zeros((P, Q, 4), dtype=torch.int32, device="cuda")
    out = torch.zeros((Q, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    bounds = [M * p // P for p in range(P + 1)]
    for p in range(P):
        m.hamming_top2_device(dq, Q, dm[bounds[p]:], bounds[p + 1] - bounds[p], bounds[p], parts[p])
    m.top2_merge_device(parts, P, Q, out)
    ctx.synchronize()
    got = out.cpu().numpy()
    for i, f in enumerate(("d1", "i1", "d2", "i2")):
        assert (got[:, i] == ref[f]).all(), f
    # the host-side merge used by the CPU-only multi-rank test must agree as well
    host = oracle.top2_merge(parts.cpu().numpy().view(oracle.TOP2).reshape(P, Q))
    for i, f in enumerate(("d1", "i1", "d2", "i2")):
        assert (host[f] == ref[f]).all(), f


def test_top2_differential_fuzz():
    """tools/fuzz_match.py: random Q / M around the tile, slice and small-query boundaries, heavy ties, index bases"""
    import os
    import subprocess
    import sys
    from util import ROOT
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_match.py"), "60", "2"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert p.returncode == 0 and " 0 mismatches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
