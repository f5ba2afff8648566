"""``gemlib.mcmc`` symbols the reference imports (mcmc_kernel_factory.py:5-9, inference.py:18-19), backed by the
CUDA library: ``UncalibratedEventTimesUpdate``, ``UncalibratedOccultUpdate``, ``TransitionTopology``,
``MultiScanKernel``, ``GibbsKernel``, ``Posterior``.

State convention of this model (inference.py:563-576): ``state = [u, events]`` -- part 0 the unconstrained
parameter block ``[B, P]``, part 1 the censored event tensor.  On the device the events live in the chain
set's caches and are updated IN PLACE by the discrete kernels; ``DeviceEvents`` is the handle that stands for
them in the state list (``.to_tensor()`` exports the reference layout ``[B, M, T, 3]`` float64).

``GibbsKernel.one_step`` has two routes that consume the same Philox random numbers and give bit-identical
chains (tests/test_gpu_kernel_tree.py):
  * composed -- every kernel of the tree issues its own C-ABI calls, in the reference's scan order;
  * fused    -- when the tree is the reference's standard one (inference.py:86-101, mcmc_kernel_factory.py:116-168)
                the whole sweep is ONE ``seir_mcmc_sweep`` call.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .. import _native as nat
from .. import tfp_mcmc as tm
from ..posterior import Posterior  # noqa: F401  (gemlib.mcmc.Posterior, inference.py:18)

TransitionTopology = namedtuple("TransitionTopology", ["prev", "target", "next"])  # mcmc_kernel_factory.py:102-104

GibbsKernelResults = namedtuple("GibbsKernelResults", ["target_log_prob", "inner_results"])
EventTimesResults = namedtuple("EventTimesResults", ["log_acceptance_correction", "target_log_prob", "m", "t", "delta_t", "x_star"])


class DeviceEvents:
    """The event tensors of B chains resident in the engine's caches (day-slab int32 layout)."""

    def __init__(self, engine, num_chains):
        self.engine = engine
        self.num_chains = int(num_chains)
        self.generation = engine.generation(self.num_chains)  # of the ingest these events came from

    def check(self, what="DeviceEvents"):
        """Raise if the chain set was re-ingested since (e.g. by ``model.log_prob(explicit events)`` with as many chains):
        the in-place sampler state this handle stands for would be gone."""
        self.engine.require_events(self.num_chains, what, self.generation)

    @property
    def shape(self):
        return (self.num_chains, self.engine.M, self.engine.T, 3)

    def to_tensor(self, dtype=torch.float64) -> torch.Tensor:
        """The events in the reference layout [B,M,T,3]: float64 (the reference's dtype) or torch.uint16 (compact)."""
        self.check("DeviceEvents.to_tensor")
        if dtype == torch.uint16:
            return self.engine.export_events_u16(self.num_chains)
        return self.engine.export_events(self.num_chains).to(dtype)

    def numpy(self):
        return self.to_tensor().cpu().numpy()


def _as_device_events(engine, events):
    if isinstance(events, DeviceEvents):
        events.check()
        return events
    B = engine.ingest(events)
    return DeviceEvents(engine, B)


class _ConditionalTarget:
    """``lambda x: target_log_prob_fn(*state with part idx replaced by x)`` (what GibbsKernel hands to each kernel
    maker [recall gemlib]), keeping a route to the engine and to the parameter block the events are conditioned on."""

    def __init__(self, parent, parts, index):
        self.parent, self.parts, self.index = parent, list(parts), index
        self.engine = tm.engine_of(parent)
        if len(parts) == 2 and index == 1:
            self.conditioned_u = parts[0]
        else:
            self.conditioned_u = getattr(parent, "conditioned_u", None)

    def __call__(self, x):
        args = list(self.parts)
        args[self.index] = x
        return self.parent(*args)


# ---- discrete kernels (a6, a7) -----------------------------------------------------------------------
class _DeviceUpdate:
    """Common part of the two uncalibrated discrete kernels: proposal on the device (``seir_propose``), then
    delta log-likelihood + MH decision + in-place commit (``seir_update_step``)."""

    is_calibrated = False
    kind = None

    def __init__(self, target_log_prob_fn, name=None):
        self.target_log_prob_fn = target_log_prob_fn
        self.engine = tm.engine_of(target_log_prob_fn)
        self.name = name
        self._local_step = 0

    @property
    def slot(self):
        return 2 * self.kind + self.update_spec.target  # accepted_results memory of the four standard kernels

    def bootstrap_results(self, events):
        ev = _as_device_events(self.engine, events)
        B, dev = ev.num_chains, self.engine.device
        u = getattr(self.target_log_prob_fn, "conditioned_u", None)
        if u is None:
            raise TypeError("the event kernels need the parameter block they are conditioned on: build them through "
                            "GibbsKernel (state = [u, events])")
        tlp = self.engine.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        cols = self.update_spec.mmax if self.kind == 0 else 1
        z = torch.zeros(B, cols, dtype=torch.int32, device=dev)
        return EventTimesResults(torch.zeros(B, dtype=torch.float64, device=dev), tlp, z, z.clone(), z.clone(), z.clone())

    def one_step(self, events, previous_kernel_results, seed=None):  # pragma: no cover - documented limitation
        raise NotImplementedError("the proposal and the Metropolis-Hastings decision are fused on the device: wrap this "
                                  "kernel in MetropolisHastings (as mcmc_kernel_factory.py:72,99 does)")

    def _mh_step(self, events, prev, seed: "tm.SeedPath", chain_offset=0):
        ev = _as_device_events(self.engine, events)
        B = ev.num_chains
        u = getattr(self.target_log_prob_fn, "conditioned_u", None)
        self.engine.prepare_theta(u)
        spec = self.update_spec
        if seed is None:
            seed = tm.SeedPath(0, self._local_step)
        self._local_step += 1
        counter = seed.sweep * 64 + seed.rep * 4 + self.slot
        prop, log_u = self.engine.propose(spec, B, seed.base, chain_offset, counter)
        tlp = prev.accepted_results.target_log_prob.clone()
        acc, trace, dbg = self.engine.update_step(spec, self.slot, prop, log_u, tlp, want_debug=True)
        cols = spec.mmax if self.kind == 0 else 1
        accepted = EventTimesResults(dbg[:, 1].clone(), tlp, trace[:, 0, :cols], trace[:, 1, :cols], trace[:, 2, :cols], trace[:, 3, :cols])
        proposed = EventTimesResults(dbg[:, 1].clone(), dbg[:, 2].clone(), prop[:, 0, :cols], prop[:, 1, :cols], prop[:, 2, :cols], prop[:, 3, :cols])
        return ev, tm.MetropolisHastingsResults(accepted, acc != 0, dbg[:, 3].clone(), proposed)


class UncalibratedEventTimesUpdate(_DeviceUpdate):
    """Move ``x_star`` events of transition ``target_event_id`` in ``mmax`` metapopulations from day ``t`` to
    ``t + delta_t`` (call site mcmc_kernel_factory.py:73-82; hyper-parameters example_config.yaml:26-28)."""

    kind = 0

    def __init__(self, target_log_prob_fn, target_event_id, prev_event_id, next_event_id, initial_state, dmax, mmax, nmax,
                 name=None):
        super().__init__(target_log_prob_fn, name)
        self.initial_state = initial_state
        self.update_spec = nat.SeirUpdateSpec(
            kind=0, target=int(target_event_id), prev=-1 if prev_event_id is None else int(prev_event_id),
            next=-1 if next_event_id is None else int(next_event_id), mmax=int(mmax), nmax=int(nmax), dmax=int(dmax), t0=0, t1=0)


class UncalibratedOccultUpdate(_DeviceUpdate):
    """Add or delete up to ``nmax`` occult events of ``topology.target`` at a (metapopulation, day) of ``t_range``
    (call site mcmc_kernel_factory.py:100-109; occult_nmax example_config.yaml:29; t_range inference.py:336-339)."""

    kind = 1

    def __init__(self, target_log_prob_fn, topology, cumulative_event_offset, nmax, t_range, name=None):
        super().__init__(target_log_prob_fn, name)
        self.topology = topology
        self.cumulative_event_offset = cumulative_event_offset
        self.update_spec = nat.SeirUpdateSpec(
            kind=1, target=int(topology.target), prev=-1 if topology.prev is None else int(topology.prev),
            next=-1 if topology.next is None else int(topology.next), mmax=1, nmax=int(nmax), dmax=0,
            t0=int(t_range[0]), t1=int(t_range[1]))


# ---- MultiScanKernel / GibbsKernel (a8) -----------------------------------------------------------------
class MultiScanKernel:
    """Apply ``kernel.one_step`` ``num_updates`` times and return the last inner results
    (mcmc_kernel_factory.py:122-124; confirmed by inference.py:276-280)."""

    is_calibrated = True

    def __init__(self, num_updates, kernel, name=None):
        self.num_updates = int(num_updates)
        self.inner_kernel = kernel
        self.name = name

    @property
    def engine(self):
        return self.inner_kernel.engine

    def bootstrap_results(self, state):
        return self.inner_kernel.bootstrap_results(state)

    def one_step(self, state, previous_kernel_results, seed=None, chain_offset=0):
        sp = tm.as_seed_path(seed)
        results = previous_kernel_results
        for rep in range(self.num_updates):
            state, results = self.inner_kernel.one_step(state, results, seed=sp._replace(rep=rep), chain_offset=chain_offset)
        return state, results


class GibbsKernel:
    """Metropolis-within-Gibbs: ``kernel_list = [(state_part_idx, make_kernel_fn)]``, each maker called as
    ``fn(conditional_target_log_prob_fn, state)`` (inference.py:97-101, mcmc_kernel_factory.py:122-166).
    The running ``target_log_prob`` is forwarded from block to block."""

    is_calibrated = True

    def __init__(self, target_log_prob_fn, kernel_list, name=None, fused=True, chain_offset=None):
        self.target_log_prob_fn = target_log_prob_fn
        self.kernel_list = list(kernel_list)
        self.name = name
        self.fused = bool(fused)
        self.engine = tm.engine_of(target_log_prob_fn)
        # global id of this rank's first chain: RNG streams are keyed by global chain id
        self.chain_offset = int(getattr(self.engine, "chain_offset", 0) if chain_offset is None else chain_offset)
        self.sweep_counter = 0

    # -- state handling --
    def normalise_state(self, state):
        """[u, events] -> [u [B,P] CUDA f64, DeviceEvents] (events are ingested into the caches once)."""
        if not isinstance(state, (list, tuple)):
            return _as_device_events(self.engine, state)
        if len(state) != 2:
            raise ValueError("state must be [unconstrained_params, events] (inference.py:563-576)")
        u = self.engine.to_device(state[0], (self.engine.P,))
        ev = _as_device_events(self.engine, state[1])
        if u.shape[0] == 1 and ev.num_chains > 1:
            u = u.expand(ev.num_chains, -1).contiguous()
        return [u, ev]

    def _kernels(self, parts):
        out = []
        for idx, maker in self.kernel_list:
            cond = _ConditionalTarget(self.target_log_prob_fn, parts, idx)
            out.append((idx, maker(cond, parts)))
        return out

    def bootstrap_results(self, state):
        parts = self.normalise_state(state)
        single = not isinstance(parts, list)
        plist = [parts] if single else parts
        inner = [k.bootstrap_results(plist[idx]) for idx, k in self._kernels(plist)]
        tlp = tm.unnest.get_innermost(inner[-1], "target_log_prob")
        return GibbsKernelResults(tlp, inner)

    # -- the reference's standard tree -> one seir_mcmc_sweep --
    def _standard_tree(self, kernels):
        if len(kernels) != 2 or kernels[0][0] != 0 or kernels[1][0] != 1:
            return None
        base, wrappers = tm.hmc_stack(kernels[0][1])
        scan = kernels[1][1]
        if base is None or not isinstance(scan, MultiScanKernel) or not isinstance(scan.inner_kernel, GibbsKernel):
            return None
        return base, wrappers, scan

    @staticmethod
    def _sweep_spec(base, scan, moves, chain_offset, seed_base):
        se, ei, ose, oei = (k.inner_kernel.update_spec for _, k in moves)
        ok = (se.kind == 0 and ei.kind == 0 and ose.kind == 1 and oei.kind == 1 and se.target == 0 and ei.target == 1 and
              ose.target == 0 and oei.target == 1 and se.prev == -1 and ose.prev == -1 and ei.prev == 0 and oei.prev == 0 and
              (se.dmax, se.nmax, se.mmax) == (ei.dmax, ei.nmax, ei.mmax) and (ose.nmax, ose.t0, ose.t1) == (oei.nmax, oei.t0, oei.t1))
        if not ok:
            return None
        return nat.SeirSweepSpec(num_leapfrog_steps=base.num_leapfrog_steps, num_event_time_updates=scan.num_updates, dmax=se.dmax,
                                 nmax=se.nmax, mmax=se.mmax, occult_nmax=ose.nmax, t0=ose.t0, t1=ose.t1, chain_offset=chain_offset,
                                 reserved=0, seed=seed_base)

    def _fused_step(self, parts, prev, sp, tree):
        base, wrappers, scan = tree
        u, ev = parts
        inner_gibbs = scan.inner_kernel
        moves = inner_gibbs._kernels([ev])
        if len(moves) != 4 or not all(isinstance(k, tm.MetropolisHastings) for _, k in moves):
            return None
        spec = self._sweep_spec(base, scan, moves, self.chain_offset, sp.base)
        if spec is None or scan.num_updates > 16:
            return None
        eng, B, dev = self.engine, ev.num_chains, self.engine.device
        hmc_prev = prev.inner_results[0]
        # what the wrappers would hand to the base kernel
        da = next((n for n in tm.unnest._walk(hmc_prev) if isinstance(n, tm.DualAveragingResults)), None)
        step = (da.new_step_size if da is not None else tm.unnest.get_innermost(hmc_prev, "step_size")).contiguous()
        md = tm.unnest.get_innermost(hmc_prev, "momentum_distribution", default=False)
        inv_mass = md.inv_mass.contiguous() if md not in (None, False) else None
        new_u = u.clone()
        tlp = prev.target_log_prob.clone()
        hmc_acc = torch.empty(B, dtype=torch.int32, device=dev)
        hmc_dbg = torch.empty(B, 4, dtype=torch.float64, device=dev)
        upd_acc = torch.empty(4, B, dtype=torch.int32, device=dev)
        upd_tlp = torch.empty(5, B, dtype=torch.float64, device=dev)
        upd_trace = torch.empty(4, B, 4, nat.MMAX, dtype=torch.int32, device=dev)
        eng.mcmc_sweep(spec, sp.sweep, new_u, step, inv_mass, tlp, hmc_acc, upd_acc, hmc_dbg=hmc_dbg, upd_tlp=upd_tlp, upd_trace=upd_trace)
        # rebuild the results tree of the composed route
        hmc_res = tm.HMCResults(hmc_acc != 0, upd_tlp[4], hmc_dbg[:, 0], step, md if md is not False else None, hmc_dbg[:, 1])
        res0 = _rewrap(hmc_prev, wrappers, hmc_res, new_u)
        inner = []
        for s, (_, k) in enumerate(moves):
            cols = k.inner_kernel.update_spec.mmax if k.inner_kernel.kind == 0 else 1
            tr = upd_trace[s]
            acc_res = EventTimesResults(None, upd_tlp[s], tr[:, 0, :cols], tr[:, 1, :cols], tr[:, 2, :cols], tr[:, 3, :cols])
            inner.append(tm.MetropolisHastingsResults(acc_res, upd_acc[s] != 0, None, None))
        res1 = GibbsKernelResults(upd_tlp[3], inner)
        return [new_u, ev], GibbsKernelResults(tlp, [res0, res1])

    # -- a whole window of the FIXED standard kernel in one seir_mcmc_burst call --
    def burst_plan(self, state):
        """(parts, tree, moves, base) when this kernel is the reference's standard tree with a plain (non-adapting) HMC kernel on
        the parameter block -- what make_fixed_window_sampler builds (inference.py:199-242) -- else None."""
        if not self.fused:
            return None
        parts = self.normalise_state(state)
        if not isinstance(parts, list):
            return None
        tree = self._standard_tree(self._kernels(parts))
        if tree is None or tree[1]:  # adaptation wrappers need the host between sweeps
            return None
        base, _, scan = tree
        moves = scan.inner_kernel._kernels([parts[1]])
        if len(moves) != 4 or not all(isinstance(k, tm.MetropolisHastings) for _, k in moves) or scan.num_updates > 16:
            return None
        return parts, tree, moves, base

    def burst(self, plan, previous_kernel_results, num_results, keep_every, sp, events_dtype=None):
        """``num_results * keep_every`` sweeps on the device in ONE C call (tfp.mcmc.sample_chain over a burst,
        inference.py:107-117, 232-240), every ``keep_every``-th kept.  Returns (u draws [n,B,P], events draws [n,B,M,T,3] or
        None, results with a leading [n] axis on every field, final state, final results); bit-identical to ``one_step`` in a
        loop (tests/test_gpu_kernel_tree.py)."""
        parts, tree, moves, base = plan
        _, _, scan = tree
        u, ev = parts
        spec = self._sweep_spec(base, scan, moves, self.chain_offset, sp.base)
        if spec is None:
            return None
        eng, B, dev = self.engine, ev.num_chains, self.engine.device
        n, e = int(num_results), int(keep_every)
        slots = n + (1 if e > 1 else 0)
        hmc_prev = previous_kernel_results.inner_results[0]
        step = tm.unnest.get_innermost(hmc_prev, "step_size").contiguous()
        md = tm.unnest.get_innermost(hmc_prev, "momentum_distribution", default=False)
        inv_mass = md.inv_mass.contiguous() if md not in (None, False) else None
        new_u = u.clone()
        tlp = previous_kernel_results.target_log_prob.clone()
        hmc_acc = torch.empty(slots, B, dtype=torch.int32, device=dev)
        hmc_dbg = torch.empty(slots, B, 4, dtype=torch.float64, device=dev)
        upd_acc = torch.empty(slots, 4, B, dtype=torch.int32, device=dev)
        upd_tlp = torch.empty(slots, 5, B, dtype=torch.float64, device=dev)
        upd_trace = torch.empty(slots, 4, B, 4, nat.MMAX, dtype=torch.int32, device=dev)
        draws = torch.empty(slots, B, eng.P, dtype=torch.float64, device=dev)
        ev16 = ovf = None
        if events_dtype is not None:
            ev16 = torch.empty(slots, B, eng.M, eng.T, 3, dtype=torch.uint16, device=dev)
            ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        eng.mcmc_burst(spec, sp.sweep, n * e, new_u, step, inv_mass, tlp, hmc_acc, upd_acc, hmc_dbg=hmc_dbg, upd_tlp=upd_tlp,
                       upd_trace=upd_trace, draws=draws, keep_every=e, events_u16=ev16, overflow=ovf)
        events = None
        if ev16 is not None:
            if int(ovf.item()) != 0:
                raise OverflowError("an event count exceeds 65535: set Mcmc.store_events_as: float64")
            events = ev16[:n] if events_dtype == torch.uint16 else ev16[:n].to(events_dtype)
        md_out = md if md is not False else None
        stepn = step.unsqueeze(0).expand(n, B)
        hmc_res = tm.HMCResults(hmc_acc[:n] != 0, upd_tlp[:n, 4], hmc_dbg[:n, :, 0], stepn, md_out, hmc_dbg[:n, :, 1])
        inner = []
        for s_, (_, k) in enumerate(moves):
            cols = k.inner_kernel.update_spec.mmax if k.inner_kernel.kind == 0 else 1
            tr = upd_trace[:n, s_]
            acc_res = EventTimesResults(None, upd_tlp[:n, s_], tr[:, :, 0, :cols], tr[:, :, 1, :cols], tr[:, :, 2, :cols], tr[:, :, 3, :cols])
            inner.append(tm.MetropolisHastingsResults(acc_res, upd_acc[:n, s_] != 0, None, None))
        stacked = GibbsKernelResults(upd_tlp[:n, 3], [hmc_res, GibbsKernelResults(upd_tlp[:n, 3], inner)])

        def last(x):  # the final kernel results: slot n - 1
            return None if x is None else x[n - 1]

        hmc_last = tm.HMCResults(last(hmc_res.is_accepted), last(hmc_res.target_log_prob), last(hmc_res.log_accept_ratio), step, md_out,
                                 last(hmc_res.proposed_target_log_prob))
        inner_last = [tm.MetropolisHastingsResults(EventTimesResults(None, *[last(f) for f in r.accepted_results[1:]]), last(r.is_accepted), None, None)
                      for r in inner]
        final = GibbsKernelResults(tlp, [hmc_last, GibbsKernelResults(upd_tlp[n - 1, 3], inner_last)])
        return draws[:n], events, stacked, [new_u, ev], final

    def one_step(self, state, previous_kernel_results, seed=None, chain_offset=None):
        parts = self.normalise_state(state)
        single = not isinstance(parts, list)
        plist = [parts] if single else list(parts)
        sp = tm.as_seed_path(seed, self.sweep_counter)
        off = self.chain_offset if chain_offset is None else chain_offset
        kernels = self._kernels(plist)
        if self.fused and not single:
            tree = self._standard_tree(kernels)
            if tree is not None:
                out = self._fused_step(plist, previous_kernel_results, sp, tree)
                if out is not None:
                    return out
        tlp = previous_kernel_results.target_log_prob
        inner = []
        for slot, (idx, _) in enumerate(self.kernel_list):
            # rebuild the kernel against the CURRENT other parts (the parameter block may have moved)
            kernel = self.kernel_list[slot][1](_ConditionalTarget(self.target_log_prob_fn, plist, idx), plist)
            prev = _forward_tlp(previous_kernel_results.inner_results[slot], tlp)
            new_part, res = kernel.one_step(plist[idx], prev, seed=sp._replace(slot=slot) if single else sp, chain_offset=off)
            plist[idx] = new_part
            tlp = tm.unnest.get_innermost(res, "target_log_prob") if not isinstance(res, GibbsKernelResults) else res.target_log_prob
            inner.append(res)
        return (plist[0] if single else plist), GibbsKernelResults(tlp, inner)


def _forward_tlp(results, tlp):
    """Put the running target log-prob into a kernel's previous results (GibbsKernel forwards it between blocks)."""
    if isinstance(results, tm.MetropolisHastingsResults):
        return results._replace(accepted_results=results.accepted_results._replace(target_log_prob=tlp))
    if isinstance(results, GibbsKernelResults):
        return results._replace(target_log_prob=tlp)
    return results  # gradient-based kernels re-bootstrap on the device (fresh value + gradient)


def _rewrap(prev, wrappers, hmc_res, new_u):
    """Apply the adaptation wrappers' post-step arithmetic, innermost first, to a base HMC result."""
    prevs = []
    node = prev
    for _ in wrappers:
        prevs.append(node)
        node = node.inner_results
    res = hmc_res
    for w, p in zip(reversed(wrappers), reversed(prevs)):
        if isinstance(w, tm.DualAveragingStepSizeAdaptation):
            res = w.adapt(p, res)
        else:
            res = w.adapt(p, res, new_u)
    return res


__all__ = ["UncalibratedEventTimesUpdate", "UncalibratedOccultUpdate", "TransitionTopology", "MultiScanKernel", "GibbsKernel",
           "GibbsKernelResults", "EventTimesResults", "DeviceEvents", "Posterior"]
