"""GPU parity of the preconditioned HMC transition (a9) against the oracle's leapfrog, RNG-free: explicit
momentum and explicit log u.  Accept/reject bit-exact, positions and log-probs to 1e-10 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,T,B,step,with_mass", [(11, 32, 3, 0.0005, False), (382, 84, 2, 0.00002, True)])
def test_hmc_transition_matches_oracle(M, T, B, step, with_mass):
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    pb = syn.make_problem(M, T, chains=B, seed=11)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    u0 = so.unconstrain(pb["theta"])
    P = u0.shape[1]
    rng = np.random.default_rng(3)
    inv_mass = np.exp(rng.normal(0.0, 0.5, size=(B, P))) if with_mass else None
    eng.ingest(pb["events"])
    n_acc = 0
    u_dev = torch.as_tensor(u0, device="cuda").clone()
    u_ora = u0.copy()
    for it in range(3):
        z = rng.normal(size=(B, P))
        mom = z / np.sqrt(inv_mass) if with_mass else z
        log_u = np.log(rng.random(B))
        steps = np.full(B, step) * (1.0 + 0.1 * np.arange(B))
        tlp, acc, dbg = eng.hmc_step(u_dev, mom, log_u, steps, inv_mass, num_leapfrog_steps=16, want_debug=True)
        tlp, acc, dbg = tlp.cpu().numpy(), acc.cpu().numpy(), dbg.cpu().numpy()
        for b in range(B):
            fn = lambda x, b=b: om.joint_log_prob_and_grad(x, pb["events"][b])
            res = so.hmc_transition(fn, u_ora[b], mom[b], log_u[b], steps[b], 16, None if inv_mass is None else inv_mass[b])
            assert bool(acc[b]) == res["is_accepted"], (it, b, dbg[b], res["log_accept_ratio"])
            if np.isfinite(res["log_accept_ratio"]):
                assert abs(dbg[b, 0] - res["log_accept_ratio"]) <= 1e-6 + 1e-9 * abs(res["proposed_tlp"]), (dbg[b], res["log_accept_ratio"])
                assert abs(dbg[b, 1] - res["proposed_tlp"]) <= 1e-10 * abs(res["proposed_tlp"])
            assert abs(tlp[b] - res["target_log_prob"]) <= 1e-10 * abs(res["target_log_prob"])
            u_ora[b] = res["state"]
            n_acc += int(res["is_accepted"])
        np.testing.assert_allclose(u_dev.cpu().numpy(), u_ora, rtol=1e-9, atol=1e-12)
    assert n_acc > 0
    eng.close()
