"""Two-phase tensor-core VQ search (csrc/vq_umma.cu: phase 1 scores all codewords with 3xTF32 tcgen05 MMAs and keeps
every candidate within a provable margin of the approximate minimum, phase 2 re-scores the candidates with the
oracle's exact sequential-fma arithmetic) against the C oracle.  Indices, quantised rows and the commitment term must
equal the exhaustive search bit for bit, including exact ties (duplicated codewords: the lowest index wins) and rows
that coincide with a codeword."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("heads,K,n", [(4, 256, 3840), (4, 256, 960), (4, 64, 3840), (4, 128, 777), (2, 64, 100),
                                       (8, 128, 77), (1, 256, 130), (4, 256, 20011)])
def test_vq_umma_bit_exact_vs_c_oracle(heads, K, n, monkeypatch):
    from msmctts._b200 import functional as Fn
    from oracle import vq as OV
    monkeypatch.setattr(Fn, "VQ_UMMA", True)
    dev = torch.device("cuda:0")
    dim = 64
    rng = np.random.default_rng(heads * 1000 + K + n)
    z = rng.standard_normal((n, heads * dim)).astype(np.float32)
    E = rng.standard_normal((heads, dim, K)).astype(np.float32)
    E[:, :, 7] = E[:, :, 3]                       # exact ties between codewords 3 and 7 -> index 3 must win
    for h in range(heads):                        # rows that coincide with a codeword / sit between two codewords
        z[0, h * dim:(h + 1) * dim] = E[h][:, 3]
        z[1, h * dim:(h + 1) * dim] = E[h][:, K - 1]
        z[2, h * dim:(h + 1) * dim] = 0.5 * (E[h][:, 10] + E[h][:, 11])
    _, q_st, diff, idx = OV.search_c(z, E)
    q, d, i = Fn.vq_quantize(torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev), heads, dim)
    assert torch.equal(i.cpu(), torch.from_numpy(idx)), "code indices must be bit-exact"
    assert torch.equal(q.cpu(), torch.from_numpy(q_st))
    assert torch.equal(d.cpu(), torch.from_numpy(diff))
