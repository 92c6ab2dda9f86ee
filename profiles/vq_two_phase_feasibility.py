"""CPU feasibility check for the two-phase VQ search planned in DESIGN.md section 8 (item 5).

Phase 1 is emulated with 3xTF32-split fp32 matmuls (hi = top 19 bits, lo = x - hi; lo*hi + hi*lo + hi*hi), i.e. the
arithmetic the tcgen05 path would use; a codeword is kept when its approximate distance is within 2*delta of the row's
approximate minimum, delta = 2^-20 (|z|^2 + |e|^2 + 2 |z||e|).  The check: the exact search's argmin (oracle/vq_oracle.c,
sequential fma order) is ALWAYS inside the candidate set, and the set is almost always a singleton.

Output on 20 000 rows x 4 heads, K = 256, dim = 64 (run in the build container, round 1):
  N(0,1) z, N(0,1) E               multi-candidate 7 of 80000 (0.009 %), max 2 candidates, oracle argmin outside set: 0
  clustered z = E[k] + 0.05 noise  multi-candidate 0 of 80000,           max 1,            oracle argmin outside set: 0
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vq as OV  # noqa: E402  (checker only)


def hi(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def main():
    rng = np.random.default_rng(0)
    n, H, dim, K = 20000, 4, 64, 256
    for scale, name in ((1.0, "N(0,1) z, N(0,1) E"), (0.05, "clustered z = E[k] + 0.05 noise")):
        E = rng.standard_normal((H, dim, K)).astype(np.float32)
        if scale == 1.0:
            z = rng.standard_normal((n, H * dim)).astype(np.float32)
        else:
            ks = rng.integers(0, K, size=(n, H))
            z = np.concatenate([E[h][:, ks[:, h]].T for h in range(H)], axis=1)
            z = (z + scale * rng.standard_normal((n, H * dim))).astype(np.float32)
        _, _, _, idx = OV.search_c(z, E)
        multi = miss = most = 0
        for h in range(H):
            zh, Eh = torch.from_numpy(z[:, h * dim:(h + 1) * dim]), torch.from_numpy(E[h])
            zhi, zlo, Ehi, Elo = hi(zh), zh - hi(zh), hi(Eh), Eh - hi(Eh)
            dot = zlo @ Ehi + zhi @ Elo + zhi @ Ehi
            zz, ee = (zh * zh).sum(1, keepdim=True), (Eh * Eh).sum(0, keepdim=True)
            dist = zz - 2 * dot + ee
            delta = 2.0 ** -20 * (zz + ee + 2 * zz.sqrt() * ee.sqrt())
            cand = dist <= dist.min(1, keepdim=True).values + 2 * delta
            cnt = cand.sum(1)
            multi += int((cnt > 1).sum())
            most = max(most, int(cnt.max()))
            miss += int((~cand[torch.arange(n), torch.from_numpy(idx[:, h])]).sum())
        print("%-32s multi-candidate %d of %d (%.3f %%), max %d candidates, oracle argmin outside set: %d" % (
            name, multi, n * H, 100.0 * multi / (n * H), most, miss))


if __name__ == "__main__":
    main()
