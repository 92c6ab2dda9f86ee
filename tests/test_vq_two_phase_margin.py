"""Design check (CPU) for the two-phase VQ search planned in DESIGN.md section 8 item 5: score all codewords with
tensor-core arithmetic, keep every codeword within a margin of the approximate minimum, re-score the survivors exactly.

Phase 1 is emulated with 3xTF32-split fp32 matmuls (hi = top 19 bits, lo = x - hi; lo*hi + hi*lo + hi*hi), the
arithmetic the tcgen05 path uses; delta = 2^-20 (|z|^2 + |e|^2 + 2 |z||e|).  Property: the exact search's argmin
(oracle/vq_oracle.c, sequential fma order, lowest index on ties) is ALWAYS inside the candidate set, and the set is
almost always a singleton -- so phase 2 costs next to nothing and the result is identical to the exhaustive search."""
import numpy as np
import pytest
import torch


def _hi(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("clustered", [False, True], ids=["gaussian", "clustered"])
def test_candidate_set_contains_exact_argmin(clustered):
    from oracle import vq as OV
    rng = np.random.default_rng(0)
    n, H, dim, K = 6000, 4, 64, 256
    E = rng.standard_normal((H, dim, K)).astype(np.float32)
    if clustered:       # rows sit next to a codeword, like a trained encoder's outputs
        ks = rng.integers(0, K, size=(n, H))
        z = np.concatenate([E[h][:, ks[:, h]].T for h in range(H)], axis=1)
        z = (z + 0.05 * rng.standard_normal((n, H * dim))).astype(np.float32)
    else:
        z = rng.standard_normal((n, H * dim)).astype(np.float32)
    _, _, _, idx = OV.search_c(z, E)
    multi = 0
    for h in range(H):
        zh, Eh = torch.from_numpy(z[:, h * dim:(h + 1) * dim]), torch.from_numpy(E[h])
        zhi, zlo, Ehi, Elo = _hi(zh), zh - _hi(zh), _hi(Eh), Eh - _hi(Eh)
        dot = zlo @ Ehi + zhi @ Elo + zhi @ Ehi
        zz, ee = (zh * zh).sum(1, keepdim=True), (Eh * Eh).sum(0, keepdim=True)
        dist = zz - 2 * dot + ee
        delta = 2.0 ** -20 * (zz + ee + 2 * zz.sqrt() * ee.sqrt())
        cand = dist <= dist.min(1, keepdim=True).values + 2 * delta
        assert bool(cand[torch.arange(n), torch.from_numpy(idx[:, h])].all()), "exact argmin outside the candidate set"
        multi += int((cand.sum(1) > 1).sum())
    assert multi <= 0.001 * n * H, "candidate sets should almost always be singletons (%d of %d are not)" % (multi, n * H)
