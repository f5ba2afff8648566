"""The int8 tensor-core contraction (tcgen05 kind::i8, error-free digit splitting, csrc/contract_i8.cu) against the FP64
DMMA kernel (csrc/contract.cu) on the same ingested events: both compute Bc = Cstar (I/N) (model_spec.py:262).
Bound asserted: |Bc_i8 - Bc_f64| <= 1e-13 * sum_j |Cs[j,i]| I_j  (the truncation of Cs to 49 bits below its column maximum
plus FP64 summation-order differences of the reference kernel)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,T,B", [(382, 84, 5), (382, 84, 64), (250, 33, 7), (128, 20, 3)])
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
    assert np.all(err <= 1e-13 * bound + 1e-300), float((err / np.maximum(bound, 1e-300)).max())
    # and against a plain numpy contraction
    exact = np.einsum("ij,bjt->bti", consts["Cstar"] / consts["N"][None, :], I)
    assert np.all(np.abs(got - exact) <= 1e-13 * bound + 1e-300)
    eng.close()
