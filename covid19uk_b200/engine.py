"""Host-side owner of the native model / chain-set handles.

PyTorch supplies device memory and streams only; all arithmetic of the hot path happens in
``libseir_b200.so`` through the C ABI of ``include/seir_b200.h``.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int32, c_void_p

import numpy as np
import torch

from . import _native as nat

STOICHIOMETRY = np.array([[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1]])  # model_spec.py:24
TIME_DELTA = 1.0  # model_spec.py:25
NU = 0.28  # model_spec.py:26
RATE_EPS = 0.000000001  # model_spec.py:266
CAR_RHO = 0.25  # model_spec.py:174


def _as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_double))


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_int32))


def prepare_constants(covariates) -> dict:
    """One-off host preparation, identical in meaning to model_spec.py:212-230 and :171-181."""
    C = np.array(covariates["C"], dtype=np.float64)
    np.fill_diagonal(C, 0.0)
    Cstar = C + C.T
    np.fill_diagonal(Cstar, -C.sum(axis=-2))
    W = np.atleast_1d(np.squeeze(_as_f64(covariates["W"])))
    N = np.atleast_1d(np.squeeze(_as_f64(covariates["N"])))
    weekday = _as_f64(covariates["weekday"])
    weekday_c = weekday - weekday.mean(axis=-1)
    log_area = np.log(_as_f64(covariates["area"]) / 100000000.0)
    log_area_c = log_area - log_area.mean()
    adj = _as_f64(covariates["adjacency"])
    precision = np.diag(adj.sum(axis=-1)) - CAR_RHO * adj
    scale = np.linalg.cholesky(np.linalg.inv(precision))
    rows, cols = np.nonzero(precision)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    indptr = np.zeros(precision.shape[0] + 1, np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    return dict(
        Cstar=_as_f64(Cstar), W=W, N=N, weekday_c=_as_f64(weekday_c), log_area_c=_as_f64(log_area_c),
        car_indptr=indptr, car_indices=cols.astype(np.int32), car_values=_as_f64(precision[rows, cols]),
        car_log_det_scale=float(np.sum(np.log(np.diag(scale)))),
    )


class SeirEngine:
    """One model on one CUDA device plus chain sets (caches) keyed by the number of chains."""

    def __init__(self, covariates, initial_state, initial_step, num_steps, device=None):
        if not torch.cuda.is_available():
            raise nat.NativeError("covid19uk_b200 needs a CUDA device: there is no CPU fallback for the hot path")
        self.lib = nat.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        consts = prepare_constants(covariates)
        init = _as_f64(initial_state)
        self.M = int(init.shape[0])
        self.T = int(num_steps)
        self.P = 6 + (self.T - 1) + self.M
        if consts["Cstar"].shape != (self.M, self.M):
            raise ValueError(f"covariates['C'] must be [{self.M},{self.M}]")
        self._keep = (consts, init)
        spec = nat.SeirSpec(
            num_meta=self.M, num_steps=self.T, initial_step=int(initial_step),
            n_commute_volume=int(consts["W"].shape[0]), n_weekday=int(consts["weekday_c"].shape[0]),
            car_nnz=int(consts["car_indices"].shape[0]), time_delta=TIME_DELTA, nu=NU, rate_eps=RATE_EPS,
            car_log_det_scale=consts["car_log_det_scale"],
            cstar=_dptr(consts["Cstar"]), population=_dptr(consts["N"]), commute_volume=_dptr(consts["W"]),
            weekday_c=_dptr(consts["weekday_c"]), log_area_c=_dptr(consts["log_area_c"]), initial_state=_dptr(init),
            car_indptr=_iptr(consts["car_indptr"]), car_indices=_iptr(consts["car_indices"]),
            car_values=_dptr(consts["car_values"]),
        )
        handle = c_void_p()
        nat.check(self.lib.seir_model_create(byref(spec), self.device.index, byref(handle)))
        self._model = handle
        self._chains: dict[int, c_void_p] = {}
        # generation of the events held by each chain set: 0 = nothing ingested yet, +1 per ingest.  The samplers' event
        # handles (gemlib.mcmc.DeviceEvents, inference.sampler.ChainSet) remember the generation they were made at, so a
        # later explicit-events evaluation with the same number of chains cannot silently replace their in-place state.
        self._generation: dict[int, int] = {}
        self.initial_state = init
        self.chain_offset = 0  # global id of this rank's first chain (multi-GPU: set by the launcher)

    # ---- lifetime ----
    def close(self):
        for h in self._chains.values():
            self.lib.seir_chains_destroy(h)
        self._chains.clear()
        if self._model is not None:
            self.lib.seir_model_destroy(self._model)
            self._model = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def chains(self, B: int) -> c_void_p:
        h = self._chains.get(B)
        if h is None:
            h = c_void_p()
            nat.check(self.lib.seir_chains_create(self._model, int(B), byref(h)))
            self._chains[B] = h
        return h

    def generation(self, B: int) -> int:
        return self._generation.get(int(B), 0)

    def _ingested(self, B: int):
        self._generation[int(B)] = self._generation.get(int(B), 0) + 1

    def require_events(self, B: int, what: str, generation=None):
        """The *_cached / sampler / analytics entry points read the caches of the last ingested event tensor."""
        g = self._generation.get(int(B), 0)
        if g == 0:
            raise RuntimeError(f"{what}: no events have been ingested for a set of {B} chains (call ingest() / log_prob(events, ...) first)")
        if generation is not None and generation != g:
            raise RuntimeError(f"{what}: the {B}-chain caches were re-ingested (generation {g}) after this handle was made "
                               f"(generation {generation}); its in-place sampler state is gone")

    def chains_bytes(self, B: int) -> int:
        return int(self.lib.seir_chains_bytes(self.chains(B)))

    # ---- helpers ----
    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, a, shape_tail):
        """numpy / torch -> contiguous float64 CUDA tensor with a leading chain axis."""
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.float64))
        t = t.to(device=self.device, dtype=torch.float64)
        if t.dim() == len(shape_tail):
            t = t.unsqueeze(0)
        if tuple(t.shape[1:]) != tuple(shape_tail):
            raise ValueError(f"expected trailing shape {tuple(shape_tail)}, got {tuple(t.shape)}")
        return t.contiguous()

    # ---- a1 ----
    def compute_state(self, events: torch.Tensor) -> torch.Tensor:
        ev = self.to_device(events, (self.M, self.T, 3))
        out = torch.empty((ev.shape[0], self.M, self.T, 4), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_compute_state(self._model, ev.shape[0], c_void_p(ev.data_ptr()), c_void_p(out.data_ptr()), self._stream()))
        return out

    # ---- a4 / a5 ----
    def ingest(self, events: torch.Tensor):
        ev = self.to_device(events, (self.M, self.T, 3))
        nat.check(self.lib.seir_ingest_events(self.chains(ev.shape[0]), c_void_p(ev.data_ptr()), self._stream()))
        self._ingested(ev.shape[0])
        return ev.shape[0]

    def log_prob_cached(self, theta, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR):
        th = self.to_device(theta, (self.P,))
        if parts & nat.PART_SEIR:
            self.require_events(th.shape[0], "log_prob_cached")
        out = torch.empty((th.shape[0],), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_log_prob_cached(self.chains(th.shape[0]), c_void_p(th.data_ptr()), kind, parts, c_void_p(out.data_ptr()), self._stream()))
        return out

    def log_prob(self, events, theta, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR, out=None):
        ev = self.to_device(events, (self.M, self.T, 3))
        th = self.to_device(theta, (self.P,))
        if th.shape[0] != ev.shape[0]:
            raise ValueError("events and theta disagree on the number of chains")
        if out is None:
            out = torch.empty((th.shape[0],), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_log_prob(self.chains(ev.shape[0]), c_void_p(ev.data_ptr()), c_void_p(th.data_ptr()), kind, parts, c_void_p(out.data_ptr()), self._stream()))
        if parts & nat.PART_SEIR:
            self._ingested(ev.shape[0])
        return out

    def value_and_grad_cached(self, theta, kind=nat.THETA_UNCONSTRAINED, parts=nat.PART_JOINT):
        th = self.to_device(theta, (self.P,))
        if parts & nat.PART_SEIR:
            self.require_events(th.shape[0], "value_and_grad_cached")
        out = torch.empty((th.shape[0],), dtype=torch.float64, device=self.device)
        grad = torch.empty_like(th)
        nat.check(self.lib.seir_log_prob_grad_cached(self.chains(th.shape[0]), c_void_p(th.data_ptr()), kind, parts, c_void_p(out.data_ptr()), c_void_p(grad.data_ptr()), self._stream()))
        return out, grad

    def log_prob_host(self, h_events: torch.Tensor, h_theta: torch.Tensor, h_out: torch.Tensor, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR):
        """Host-buffer entry point: (pinned) CPU tensors in, CPU tensor out, synchronous."""
        B = h_events.shape[0]
        assert h_events.dtype == torch.float64 and h_events.is_contiguous() and not h_events.is_cuda
        nat.check(self.lib.seir_log_prob_host(self.chains(B), c_void_p(h_events.data_ptr()), c_void_p(h_theta.data_ptr()), kind, parts, c_void_p(h_out.data_ptr())))
        if parts & nat.PART_SEIR:
            self._ingested(B)
        return h_out

    def last_h2d_bytes(self, B: int) -> int:
        """Bytes the last log_prob_host call moved host->device (events as shipped + theta)."""
        return int(self.lib.seir_last_h2d_bytes(self.chains(B)))

    def run_stage(self, B, stage, events=None, theta=None, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR, out=None, grad=None):
        """Enqueue one kernel of the pipeline (measurement hook, see seir_run_stage)."""
        ptr = lambda t: c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
        if stage == 0:
            self._ingested(B)
        nat.check(self.lib.seir_run_stage(self.chains(B), stage, ptr(events), ptr(theta), kind, parts, ptr(out), ptr(grad), self._stream()))

    # ---- a6 / a7: discrete updates ----
    def prepare_theta(self, theta, kind=nat.THETA_UNCONSTRAINED):
        """Load the rate factors of `theta` [B,P] for the discrete updates (after any change of theta)."""
        th = self.to_device(theta, (self.P,))
        self.require_events(th.shape[0], "prepare_theta")
        nat.check(self.lib.seir_prepare_theta(self.chains(th.shape[0]), c_void_p(th.data_ptr()), kind, self._stream()))
        return th

    def update_step(self, spec: "nat.SeirUpdateSpec", slot, proposal, log_u, tlp, want_debug=False):
        """One MH step of a discrete kernel for every chain (explicit proposals, explicit log u).

        proposal int32 [B,4,MMAX] (rows m, t, delta_t, x_star), log_u [B], tlp [B] (updated in place).
        Returns (is_accepted [B] int32, trace [B,4,MMAX] int32, debug [B,4] or None)."""
        B = tlp.shape[0]
        self.require_events(B, "update_step")
        prop = torch.as_tensor(proposal, dtype=torch.int32, device=self.device).contiguous()
        lu = torch.as_tensor(log_u, dtype=torch.float64, device=self.device).contiguous()
        assert tuple(prop.shape) == (B, 4, nat.MMAX) and tlp.is_cuda and tlp.dtype == torch.float64
        acc = torch.empty((B,), dtype=torch.int32, device=self.device)
        trace = torch.empty((B, 4, nat.MMAX), dtype=torch.int32, device=self.device)
        dbg = torch.empty((B, 4), dtype=torch.float64, device=self.device) if want_debug else None
        nat.check(self.lib.seir_update_step(
            self.chains(B), byref(spec), int(slot), c_void_p(prop.data_ptr()), c_void_p(lu.data_ptr()), c_void_p(tlp.data_ptr()),
            c_void_p(acc.data_ptr()), c_void_p(trace.data_ptr()), c_void_p(dbg.data_ptr()) if dbg is not None else c_void_p(0),
            self._stream()))
        return acc, trace, dbg

    # ---- a9: HMC ----
    def hmc_step(self, u, momentum, log_u, step_size, inv_mass=None, num_leapfrog_steps=16, want_debug=False):
        """One PreconditionedHMC transition per chain (explicit momentum and log u).  `u` [B,P] is a CUDA
        float64 tensor updated in place.  Returns (tlp [B], is_accepted [B] int32, debug [B,4] or None)."""
        assert u.is_cuda and u.dtype == torch.float64 and u.is_contiguous() and u.dim() == 2 and u.shape[1] == self.P
        B = u.shape[0]
        self.require_events(B, "hmc_step")
        dev = lambda a: torch.as_tensor(a, dtype=torch.float64, device=self.device).contiguous()
        mom, lu = dev(momentum), dev(log_u)
        st = dev(step_size).expand(B).contiguous() if dev(step_size).dim() == 0 else dev(step_size)
        im = dev(inv_mass).expand(B, self.P).contiguous() if inv_mass is not None else None
        tlp = torch.empty((B,), dtype=torch.float64, device=self.device)
        acc = torch.empty((B,), dtype=torch.int32, device=self.device)
        dbg = torch.empty((B, 4), dtype=torch.float64, device=self.device) if want_debug else None
        nat.check(self.lib.seir_hmc_step(
            self.chains(B), c_void_p(u.data_ptr()), c_void_p(mom.data_ptr()), c_void_p(lu.data_ptr()), c_void_p(st.data_ptr()),
            c_void_p(im.data_ptr()) if im is not None else c_void_p(0), int(num_leapfrog_steps), c_void_p(tlp.data_ptr()),
            c_void_p(acc.data_ptr()), c_void_p(dbg.data_ptr()) if dbg is not None else c_void_p(0), self._stream()))
        return tlp, acc, dbg

    def hmc_draw(self, B, seed, chain_offset, sweep_index, inv_mass=None):
        """Momentum [B,P] ~ N(0, diag(1/inv_mass)) and log u [B] from the Philox streams of the fused sweep."""
        mom = torch.empty((B, self.P), dtype=torch.float64, device=self.device)
        lu = torch.empty((B,), dtype=torch.float64, device=self.device)
        im = None
        if inv_mass is not None:
            im = torch.as_tensor(inv_mass, dtype=torch.float64, device=self.device).expand(B, self.P).contiguous()
        nat.check(self.lib.seir_hmc_draw(self.chains(B), int(seed), int(chain_offset), int(sweep_index),
                                         c_void_p(im.data_ptr()) if im is not None else c_void_p(0),
                                         c_void_p(mom.data_ptr()), c_void_p(lu.data_ptr()), self._stream()))
        return mom, lu

    # ---- device-side proposals and the fused sweep (a8) ----
    def propose(self, spec: "nat.SeirUpdateSpec", B, seed, chain_offset, counter):
        """Draw one proposal per chain on the device -> (proposal int32 [B,4,MMAX], log_u [B])."""
        self.require_events(B, "propose")
        prop = torch.empty((B, 4, nat.MMAX), dtype=torch.int32, device=self.device)
        lu = torch.empty((B,), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_propose(self.chains(B), byref(spec), int(seed), int(chain_offset), int(counter),
                                        c_void_p(prop.data_ptr()), c_void_p(lu.data_ptr()), self._stream()))
        return prop, lu

    def mcmc_sweep(self, spec: "nat.SeirSweepSpec", sweep_index, u, step_size, inv_mass, tlp, hmc_accept, upd_accept,
                   hmc_dbg=None, upd_tlp=None, upd_trace=None):
        """One Metropolis-within-Gibbs sweep for every chain; all tensors are caller-owned CUDA tensors."""
        B = u.shape[0]
        self.require_events(B, "mcmc_sweep")
        p = lambda t: c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
        nat.check(self.lib.seir_mcmc_sweep(self.chains(B), byref(spec), int(sweep_index), p(u), p(step_size), p(inv_mass), p(tlp),
                                           p(hmc_accept), p(hmc_dbg), p(upd_accept), p(upd_tlp), p(upd_trace), self._stream()))

    def mcmc_burst(self, spec: "nat.SeirSweepSpec", sweep_index0, num_sweeps, u, step_size, inv_mass, tlp, hmc_accept, upd_accept,
                   hmc_dbg=None, upd_tlp=None, upd_trace=None, draws=None, keep_every=1, events_u16=None, overflow=None):
        """``num_sweeps`` sweeps with a fixed step size / mass matrix in one call (seir_mcmc_burst): the result tensors carry
        a leading [num_sweeps] axis; chains and traces are bit-identical to ``num_sweeps`` calls of :meth:`mcmc_sweep`.
        ``keep_every = e`` keeps every e-th sweep only (result tensors: ``num_sweeps // e`` kept slots + one scratch slot);
        ``events_u16`` [slots,B,M,T,3] (torch.uint16) receives the compact events of every kept sweep."""
        B = u.shape[0]
        self.require_events(B, "mcmc_burst")
        p = lambda t: c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
        nat.check(self.lib.seir_mcmc_burst(self.chains(B), byref(spec), int(sweep_index0), int(num_sweeps), p(u), p(step_size), p(inv_mass),
                                           p(tlp), p(hmc_accept), p(hmc_dbg), p(upd_accept), p(upd_tlp), p(upd_trace), p(draws),
                                           int(keep_every), p(events_u16), p(overflow), self._stream()))

    def export_events_u16(self, B: int, out=None) -> torch.Tensor:
        """The current events as uint16 counts [B,M,T,3] (compact posterior storage); raises if a count exceeds 65535."""
        self.require_events(B, "export_events_u16")
        if out is None:
            out = torch.empty((B, self.M, self.T, 3), dtype=torch.uint16, device=self.device)
        ovf = torch.zeros(1, dtype=torch.int32, device=self.device)
        nat.check(self.lib.seir_export_events_u16(self.chains(B), c_void_p(out.data_ptr()), c_void_p(ovf.data_ptr()), self._stream()))
        if int(ovf.item()) != 0:
            raise OverflowError("an event count exceeds 65535: store the events as float64 (events_dtype=torch.float64)")
        return out

    def log_prob_host_u16(self, h_events: torch.Tensor, h_theta: torch.Tensor, h_out: torch.Tensor, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR):
        """Integer host contract: (pinned) uint16 events [B,M,T,3] in, log-prob [B] out, synchronous."""
        B = h_events.shape[0]
        assert h_events.dtype in (torch.uint16, torch.int16) and h_events.is_contiguous() and not h_events.is_cuda
        nat.check(self.lib.seir_log_prob_host_u16(self.chains(B), c_void_p(h_events.data_ptr()), c_void_p(h_theta.data_ptr()), kind, parts, c_void_p(h_out.data_ptr())))
        if parts & nat.PART_SEIR:
            self._ingested(B)
        return h_out

    def export_events(self, B: int) -> torch.Tensor:
        self.require_events(B, "export_events")
        out = torch.empty((B, self.M, self.T, 3), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_export_events(self.chains(B), c_void_p(out.data_ptr()), self._stream()))
        return out

    # ---- f4: forward simulation (model.sample) ----
    def simulate(self, alpha_path, scalars, spatial_effect, initial_state, seed=0, chain_offset=0) -> torch.Tensor:
        """Chain-binomial simulation of B samples over this model's steps -> events [B,M,T,3] (CUDA float64).
        alpha_path [B,T] resolved per-step log-rate offsets, scalars [B,5] (psi, sigma_space, beta_area, gamma0, gamma1),
        spatial_effect [B,M], initial_state [B,M,4]."""
        dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64))).to(
            device=self.device, dtype=torch.float64).contiguous()
        ap, sc, sp, st = dev(alpha_path), dev(scalars), dev(spatial_effect), dev(initial_state)
        B = ap.shape[0]
        assert tuple(ap.shape) == (B, self.T) and tuple(sc.shape) == (B, 5) and tuple(sp.shape) == (B, self.M)
        assert tuple(st.shape) == (B, self.M, 4)
        out = torch.empty((B, self.M, self.T, 3), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_simulate(self._model, B, int(seed), int(chain_offset), c_void_p(ap.data_ptr()), c_void_p(sc.data_ptr()),
                                         c_void_p(sp.data_ptr()), c_void_p(st.data_ptr()), c_void_p(out.data_ptr()), self._stream()))
        return out

    # ---- f4: posterior analytics on the cached state of the ingested events ----
    def reproduction_number(self, theta) -> torch.Tensor:
        """R_it [B,T,M] (posterior/reproduction_number.py:13-45); `theta` [B,P] CONSTRAINED; events ingested before."""
        th = self.to_device(theta, (self.P,))
        B = th.shape[0]
        self.require_events(B, "reproduction_number")
        out = torch.empty((B, self.T, self.M), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_reproduction_number(self.chains(B), c_void_p(th.data_ptr()), c_void_p(out.data_ptr()), self._stream()))
        return out

    def pressure_components(self, theta):
        """(within, between) [B,M] infection pressure at the final state (posterior/within_between.py:13-56)."""
        th = self.to_device(theta, (self.P,))
        B = th.shape[0]
        self.require_events(B, "pressure_components")
        within = torch.empty((B, self.M), dtype=torch.float64, device=self.device)
        between = torch.empty_like(within)
        nat.check(self.lib.seir_pressure_components(self.chains(B), c_void_p(th.data_ptr()), c_void_p(within.data_ptr()),
                                                    c_void_p(between.data_ptr()), self._stream()))
        return within, between

    def export_contraction(self, B: int) -> torch.Tensor:
        """The cached contraction Cstar (I/N) of the ingested events, [B,T,M]."""
        self.require_events(B, "export_contraction")
        out = torch.empty((B, self.T, self.M), dtype=torch.float64, device=self.device)
        nat.check(self.lib.seir_export_contraction(self.chains(B), c_void_p(out.data_ptr()), self._stream()))
        return out

    def chain_flags(self, B: int) -> torch.Tensor:
        out = torch.empty((B,), dtype=torch.int32, device=self.device)
        nat.check(self.lib.seir_chain_flags(self.chains(B), c_void_p(out.data_ptr()), self._stream()))
        return out
