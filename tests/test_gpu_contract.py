"""The int8 tensor-core contraction (tcgen05 kind::i8, error-free digit splitting, csrc/contract_i8.cu) against the FP64
DMMA kernel (csrc/contract.cu) on the same ingested events: both compute Bc = Cstar (I/N) (model_spec.py:262).
Bound asserted: |Bc_i8 - Bc_f64| <= 1e-12 * sum_j |Cs[j,i]| I_j.  The integer kernel is exact up to the truncation of Cs to a
48-bit fixed-point value whose scale 2^e_i exceeds 4x the column maximum: per element <= 2^-47 of the column maximum
(measured worst case of the normalised error at the UK shape: 2.8e-13); on the log-probability that is < 1e-12 relative,
and tests/test_gpu_logprob.py holds its 1e-10 bound with this kernel as the default."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


# (M > 384: the long-K kernel -- K blocks outermost, 128 x 64 tiles, eight accumulator groups live in TMEM)
@pytest.mark.parametrize("M,T,B", [(382, 84, 5), (382, 84, 64), (250, 33, 7), (128, 20, 3), (500, 40, 6), (640, 21, 13), (1000, 12, 3), (2000, 9, 2)])
def test_int8_contraction_matches_fp64(M, T, B):
    import torch
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine, prepare_constants

    pb = syn.make_problem(M, T, chains=B, seed=3, distinct=min(B, 4))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    ev = torch.from_numpy(pb["events"]).cuda()
    eng.ingest(ev)
    eng.run_stage(B, 9)  # FP64 DMMA
    ref = eng.export_contraction(B).cpu().numpy()
    eng.run_stage(B, 8)  # int8 tcgen05
    got = eng.export_contraction(B).cpu().numpy()
    consts = prepare_constants(pb["covariates"])
    cs_abs = np.abs(consts["Cstar"] / consts["N"][None, :])          # |Cstar[i,j]| / N[j]
    from oracle import seir_oracle as so

    I = so.compute_state(pb["initial_state"], pb["events"])[..., 2]   # [B, M, T]
    bound = np.einsum("ij,bjt->bti", cs_abs, I)                       # sum_j |Cs| I_j
    err = np.abs(got - ref)
    assert np.all(err <= 1e-12 * bound + 1e-300), float((err / np.maximum(bound, 1e-300)).max())
    # and against a plain numpy contraction
    exact = np.einsum("ij,bjt->bti", consts["Cstar"] / consts["N"][None, :], I)
    assert np.all(np.abs(got - exact) <= 1e-12 * bound + 1e-300)
    eng.close()
