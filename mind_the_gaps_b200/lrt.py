"""Protassov likelihood-ratio test on the GPU: simulate -> fit null -> fit alternative, batched.

The reference has no library function for this; it is notebook cells
(``/root/reference/docs/notebooks/tutorial_ppp.ipynb`` cells 11-15): draw posterior samples of the null
model, ``generate_from_posteriors``, then for every simulated lightcurve fit both models (L-BFGS-B
followed by a short MCMC whose maximum log-posterior is kept) and form
``T = -2 (L_null - L_alt)``; ``p = 1 - percentileofscore(T_dist, T_obs) / 100``.

Here every stage stays on the device for a whole shard of simulations: ``mtg_sim_generate_dev`` ->
``mtg_sim_add_noise_dev`` -> ``mtg_prepare_data_dev`` -> ``mtg_lbfgs_dev`` (null, multi-start) ->
``mtg_lbfgs_dev`` (alternative, multi-start, one start nested at the null optimum so T >= 0) ->
``mtg_group_best_dev``.  The per-simulation MCMC refinement of the reference is replaced by
multi-start MLE (north_star), so agreement with the reference is statistical (distribution of T).
Simulations are independent: with ``torch.distributed`` initialised they are sharded over the ranks
(one GPU each) and only ``(T, L_null, L_alt, status)`` are gathered (NCCL all-gather).
"""
import ctypes

import numpy as np

from . import _native
from .celerite import GP
from .stats import lrt_pvalue, lrt_statistic

__all__ = ["LRTResult", "simulate_and_fit", "shard_range", "alternative_starts", "infer_null_slots", "all_gather_rows", "all_gather_varlen"]


class LRTResult(object):
    def __init__(self, T, loglike_null, loglike_alt, params_null, params_alt, status_null, status_alt, timings=None):
        self.T = T; self.loglike_null = loglike_null; self.loglike_alt = loglike_alt
        self.params_null = params_null; self.params_alt = params_alt
        self.status_null = status_null; self.status_alt = status_alt
        self.timings = timings or {}

    @property
    def n_used(self):
        """Number of simulations with a finite statistic, i.e. what ``pvalue`` is computed from."""
        return int(np.count_nonzero(np.isfinite(self.T)))

    @property
    def n_invalid(self):
        """Simulations discarded because the noise model could not be applied (negative expected counts: the reference's
        ``np.random.poisson`` raises ``ValueError("lam < 0")`` there, noise_models.py:72) or because a fit failed."""
        return int(len(self.T) - self.n_used)

    def pvalue(self, T_obs, return_n=False):
        """``1 - percentileofscore(T_dist, T_obs) / 100`` over the ``n_used`` finite statistics;
        ``return_n=True`` returns ``(p, n_used)``."""
        p = lrt_pvalue(self.T, T_obs)
        return (p, self.n_used) if return_n else p

    _FIELDS = ("T", "loglike_null", "loglike_alt", "params_null", "params_alt", "status_null", "status_alt")

    def save(self, path):
        """Write the arrays of the test to ``path`` (``.npz``): the reference's workflow keeps such per-run result files
        (``docs/workflow.md:44-45, 77, 92``: ``fits_results.dat``) so that the 10^4 simulations need not be redone."""
        np.savez_compressed(path, **{k: np.asarray(getattr(self, k)) for k in self._FIELDS})

    @classmethod
    def load(cls, path):
        with np.load(path) as z:
            return cls(*[z[k] for k in cls._FIELDS])


def shard_range(n, rank, world):
    """Contiguous block of ``n`` items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_rows(dist, local_rows, n_total, device=None):
    """Gather the per-rank row blocks of a ``shard_range`` partition into the full ``(n_total, width)``
    array on every rank (one NCCL / gloo all-gather of equal-sized, NaN-padded blocks)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(b - a for a, b in counts)
    local_rows = np.atleast_2d(np.asarray(local_rows, dtype=np.float64))
    width = local_rows.shape[1]
    if local_rows.shape[0] != counts[rank][1] - counts[rank][0]:
        raise ValueError("rank %d holds %d rows, its shard has %d" % (rank, local_rows.shape[0], counts[rank][1] - counts[rank][0]))
    pad = np.full((nmax, width), np.nan)
    pad[:local_rows.shape[0]] = local_rows
    send = torch.from_numpy(pad)
    if device is not None:
        send = send.to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    return np.concatenate([recv[r].cpu().numpy()[: counts[r][1] - counts[r][0]] for r in range(world)], axis=0)


def all_gather_varlen(dist, local_rows, device=None):
    """Concatenate row blocks of DIFFERENT lengths over the ranks, in rank order, on every rank (two all-gathers: the
    lengths, then NaN-padded blocks).  Used to pool the posterior draws of independent replica chains: the ensemble
    of one lightcurve does not shard, so every GPU runs its own chain and only the draws travel."""
    import torch
    world = dist.get_world_size()
    local_rows = np.ascontiguousarray(np.atleast_2d(np.asarray(local_rows, dtype=np.float64)))
    width = local_rows.shape[1]
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64)
    if device is not None:
        n_local = n_local.to(device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    pad = np.full((nmax, width), np.nan)
    pad[:local_rows.shape[0]] = local_rows
    send = torch.from_numpy(pad)
    if device is not None:
        send = send.to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    return np.concatenate([recv[r].cpu().numpy()[:counts[r]] for r in range(world)], axis=0), counts


def _bounds_arrays(gp):
    b = gp.get_parameter_bounds()
    lo = np.array([-np.inf if x[0] is None else x[0] for x in b], dtype=np.float64)
    hi = np.array([np.inf if x[1] is None else x[1] for x in b], dtype=np.float64)
    return lo, hi


def _term_signature(term):
    """What makes two kernel terms "the same component" for nesting: class, parameter names and frozen mask."""
    return (type(term), tuple(term.get_parameter_names(include_frozen=True)), tuple(bool(m) for m in np.atleast_1d(term.unfrozen_mask)))


def infer_null_slots(null_kernel, alt_kernel):
    """Positions, in the alternative kernel's free-parameter vector, of the parameters it shares with the null kernel.

    The match is on TERM STRUCTURE (class, parameter names, frozen mask of every term), not on parameter names: the
    null kernel's term list must appear as a contiguous run of the alternative's (``alt = extra + null`` as in the
    reference's tutorial, ``ComplexTerm + RealTerm`` vs ``RealTerm``, docs/notebooks/tutorial_ppp.ipynb cells 7/9, or
    ``alt = null + extra``).  Several matching runs (``alt = DRW + DRW``): the trailing one is taken when the
    alternative ends with the null model, anything else is ambiguous.  No match or an ambiguous one raises
    ``ValueError``: pass ``null_slots`` explicitly."""
    t0 = [_term_signature(t) for t in null_kernel.terms]
    t1 = [_term_signature(t) for t in alt_kernel.terms]
    n0, n1 = len(t0), len(t1)
    runs = [i for i in range(n1 - n0 + 1) if t1[i:i + n0] == t0]
    if not runs:
        raise ValueError("the null kernel's terms are not a contiguous run of the alternative's: pass null_slots explicitly")
    if len(runs) > 1 and runs[-1] != n1 - n0:
        raise ValueError("the null kernel matches the alternative's terms at positions %s: pass null_slots explicitly" % runs)
    start = runs[-1]
    free_per_term = [int(np.sum(np.atleast_1d(t.unfrozen_mask))) for t in alt_kernel.terms]
    off = int(sum(free_per_term[:start]))
    return list(range(off, off + int(sum(free_per_term[start:start + n0]))))


def alternative_starts(x_alt0, x_null_best, null_slots, lo, hi, n_starts, rng):
    """Start points of the alternative fits, ``(L, n_starts, nd)``: start 0 is the alternative's initial
    vector with its null-model components replaced by the null optimum (nested start: the fit can only
    improve on the null likelihood up to the cost of the extra term), the next third are Gaussian balls
    around it (10 % as ``spread_walkers``), the rest are uniform inside the finite bounds for the
    parameters that do not belong to the null model."""
    L = x_null_best.shape[0]; nd = len(x_alt0)
    base = np.tile(np.asarray(x_alt0, dtype=np.float64), (L, 1))
    base[:, null_slots] = x_null_best
    out = np.repeat(base[:, None, :], n_starts, axis=1)
    if n_starts > 1:
        n_ball = max(1, (n_starts - 1) // 3)
        out[:, 1:1 + n_ball, :] += rng.normal(size=(L, n_ball, nd)) * (0.1 * np.abs(base))[:, None, :]
        extra = [k for k in range(nd) if k not in set(null_slots)]
        n_uni = n_starts - 1 - n_ball
        for k in extra:
            if n_uni > 0 and np.isfinite(lo[k]) and np.isfinite(hi[k]):
                out[:, 1 + n_ball:, k] = rng.uniform(lo[k], hi[k], size=(L, n_uni))
            elif n_uni > 0:
                out[:, 1 + n_ball:, k] += rng.normal(size=(L, n_uni)) * (0.5 * np.abs(base[:, k]) + 0.5)[:, None]
    return np.clip(out, lo, hi)


def default_chunk(n_local, pipeline=3, refine=False):
    """Simulations per chunk of ``simulate_and_fit`` for a shard of ``n_local`` simulations.

    One chunk for up to 16384 local simulations: since the fused TK95 kernels (csrc/mtg_tk95.cu) the three stages of a
    10^4-simulation shard are 0.2 / 0.015 / 0.26 s of device work that each fill the device, and chunks in flight on several
    streams only add latency-bound tails and host-thread jitter (measured 0.45-0.60 s in one chunk against 0.48-0.90 s in
    three on two boxes; the reference grid, simulation-bound, is 2.0 s either way).  The MCMC refinement keeps the pipeline
    above two waves of ensembles (8192 simulations): its short launches pack better from several streams."""
    n_local = int(n_local)
    if refine and n_local > 8192:
        return min(8192, max(2048, -(-n_local // max(1, int(pipeline)))))
    return min(max(n_local, 1), 16384)


def simulate_and_fit(lightcurve, null_kernel, alt_kernel, null_samples, null_slots=None, pdf="Gaussian",
                     extension_factor=2, sigma_noise=None, n_starts_null=2, n_starts_alt=8, seed=0,
                     max_iter=200, device=None, distributed=None, chunk=None, pipeline=3, return_lightcurves=False,
                     simulator_kwargs=None, refine_walkers=0, refine_steps=0, refine_block=25):
    """Simulate one lightcurve per row of ``null_samples`` (posterior draws of the null kernel's
    parameters) on the window of ``lightcurve`` and MLE-fit each with both kernels.

    ``null_slots``: positions in the alternative's parameter vector of the parameters it shares with the
    null model (default: the trailing ``len(null)`` ones when the alternative is ``extra + null`` as in the
    tutorial, else the leading ones if names match).  The local simulations are cut into chunks of ``chunk``
    (default: one chunk of up to 16384 simulations; with the MCMC refinement and more than 8192 local simulations
    ``ceil(local / pipeline)`` clamped to 2048..8192) and up to ``pipeline`` chunks are in flight at once, each on its own
    CUDA stream with its own simulator workspace.

    ``simulator_kwargs``: passed to ``Simulator``; by default the regular grid is lengthened to the next 7-smooth length
    (``fft_length="fast"``: only binned lightcurves are used here, and the simulated process is statistically the same),
    ``{"fft_length": "exact"}`` keeps the reference's own grid length (usually a Bluestein FFT, ~2x slower simulation).

    ``refine_walkers, refine_steps``: when both are positive every fit is followed, as in the reference's notebook
    (``GPModelling(lc, kernel).derive_posteriors(fit=True, walkers=30, max_steps=500)`` then ``max_loglikelihood``,
    tutorial_ppp.ipynb cells 11-13), by a short stretch-move chain started in a 10 % Gaussian ball around the optimum
    (``spread_walkers``) -- all simulations of a chunk in lock step in the device ensemble sampler (kernel K8), the run cut
    into launches of ``refine_block`` steps (the same chain however it is cut) -- and the largest log-likelihood visited
    replaces the optimiser's.  Returns an ``LRTResult`` holding the arrays for ALL simulations (gathered
    over ranks when ``torch.distributed`` is initialised)."""
    import concurrent.futures
    import torch
    from .fitting import multistart_points
    lib = _native.lib()
    _native.require_device()
    dist = None
    rank, world = 0, 1
    if distributed is None:
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized():
            dist = tdist
    elif distributed:
        import torch.distributed as tdist
        dist = tdist
    if dist is not None:
        rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    dev_index = device.index if device.index is not None else 0
    _native.check(lib.mtg_set_device(dev_index))

    null_samples = np.ascontiguousarray(np.atleast_2d(null_samples), dtype=np.float64)
    n_total = null_samples.shape[0]
    i0, i1 = shard_range(n_total, rank, world)
    mine = null_samples[i0:i1]
    L = mine.shape[0]

    gp0 = GP(null_kernel, mean=0.0); gp1 = GP(alt_kernel, mean=0.0)
    m0 = gp0.device_model(); m1 = gp1.device_model()
    nd0, nd1 = m0.n_free, m1.n_free
    lo0, hi0 = _bounds_arrays(gp0); lo1, hi1 = _bounds_arrays(gp1)
    if null_slots is None:
        null_slots = infer_null_slots(null_kernel, alt_kernel)
    null_slots = [int(k) for k in null_slots]
    if len(null_slots) != nd0 or len(set(null_slots)) != nd0 or min(null_slots) < 0 or max(null_slots) >= nd1:
        raise ValueError("null_slots must name %d distinct positions of the alternative's %d free parameters" % (nd0, nd1))
    x_alt0 = gp1.get_parameter_vector().copy()

    N = lightcurve.n
    t_dev = torch.from_numpy(np.ascontiguousarray(lightcurve.times, dtype=np.float64)).to(device)
    opts = _native.MtgLbfgsOpts(int(max_iter), 10, 2.220446049250313e-09, 1e-5, 20, 0)
    pipeline = max(1, int(pipeline))
    if chunk is None:
        chunk = default_chunk(L, pipeline, int(refine_walkers) > 0 and int(refine_steps) > 0)
    chunk = max(1, int(chunk))
    starts = list(range(0, L, chunk))

    out_T = np.empty(L); out_l0 = np.empty(L); out_l1 = np.empty(L)
    out_p0 = np.empty((L, nd0)); out_p1 = np.empty((L, nd1)); out_s0 = np.empty(L, dtype=np.int32); out_s1 = np.empty(L, dtype=np.int32)
    kept = {}
    times = {"simulate_ms": 0.0, "fit_null_ms": 0.0, "fit_alt_ms": 0.0, "n_invalid": 0}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    main_stream = torch.cuda.current_stream(device)

    refine = int(refine_walkers) > 0 and int(refine_steps) > 0
    if refine and not (gp0.device_prior_supported() and gp1.device_prior_supported()):
        raise NotImplementedError("MCMC refinement runs on the device sampler, which does not restate a user-defined log_prior")

    def mcmc_refine(model, nd, lo, hi, x_best, f_best, Lc, y, diag, means, rng, sp, purpose, first):
        """max over a short ensemble run around ``x_best`` (Lc, nd); returns (loglike (Lc,), params (Lc, nd)).
        (Measured, round 2: running the null model's chain on a second stream beside the alternative's does not pay -- two
        co-resident K8 launches of 1250 ensembles take 0.84 s against 0.58 s one after the other, and started earlier the
        null chain blocks the persistent L-BFGS kernel of the alternative fits, which needs whole register files.)
        ``first``: global index of the chunk's first simulation -- folded into the Philox key, because the counters only
        carry the ensemble index INSIDE the launch and simulation k of every chunk / rank would otherwise see the same
        split, stretch, partner and accept draws."""
        ens_seed = (int(seed) * 7919 + purpose + 1000003 * (2 * int(first) + 1)) & 0xFFFFFFFFFFFFFFFF
        W = int(refine_walkers)
        std = np.abs(x_best) * 0.1
        p0 = x_best[:, None, :] + rng.normal(size=(Lc, W, nd)) * std[:, None, :]
        p0[:, 0, :] = x_best                                   # keep the optimum itself in the ensemble
        p0 = np.clip(p0, lo, hi)
        pos = torch.from_numpy(np.ascontiguousarray(p0)).to(device)
        lpw = torch.empty((Lc, W), dtype=torch.float64, device=device)
        blp = torch.empty(Lc, dtype=torch.float64, device=device); bpos = torch.empty((Lc, nd), dtype=torch.float64, device=device)
        # The run is cut into launches of `refine_block` steps (same chain: the Philox counters carry the absolute step, the
        # best-so-far state continues across launches).  One launch of all the steps keeps a CTA alive for the whole run, so
        # a shard of slightly more ensembles than the device holds at once (2 x 2368 at 30 walkers) pays a second full
        # wave for the remainder; short launches from the pipeline's streams let the block scheduler pack the remainder
        # beside the other chunks' work instead.
        done = 0
        while done < int(refine_steps):
            n_st = min(int(refine_block), int(refine_steps) - done)
            _native.check(lib.mtg_ensemble_dev(ctypes.byref(model), N, Lc, Lc, W, t_dev.data_ptr(), y.data_ptr(), diag.data_ptr(), means.data_ptr(),
                                               pos.data_ptr(), lpw.data_ptr(), 1 if done == 0 else 0, n_st, 2.0, ens_seed, done,
                                               None, None, None, blp.data_ptr(), bpos.data_ptr(), _native.FLAG_DS_MEAN, sp))
            done += n_st
        b = blp.cpu().numpy(); bp = bpos.cpu().numpy()
        better = b > f_best
        return np.where(better, b, f_best), np.where(better[:, None], bp, x_best)

    def worker(slot_chunks):
        """One pipeline slot: its own stream and simulator workspace, chunks processed in order."""
        torch.cuda.set_device(device)
        _native.check(lib.mtg_set_device(dev_index))
        stream = torch.cuda.Stream(device) if len(starts) > 1 else main_stream
        stream.wait_stream(main_stream)
        sp = stream.cuda_stream
        sim = lightcurve.get_simulator(null_kernel.get_psd, pdf, sigma_noise=sigma_noise, extension_factor=extension_factor,
                                       **dict({"fft_length": "fast"}, **(simulator_kwargs or {})))
        h = sim._sim()
        local = {"simulate_ms": 0.0, "fit_null_ms": 0.0, "fit_alt_ms": 0.0, "n_invalid": 0}
        try:
            with torch.cuda.stream(stream):
                for c0 in slot_chunks:
                    Lc = min(chunk, L - c0)
                    rng = np.random.default_rng([int(seed), int(rank), int(c0)])
                    rng_ref = np.random.default_rng([int(seed), int(rank), int(c0), 99])      # refinement draws: fits stay identical with / without it
                    p_dev = torch.from_numpy(np.ascontiguousarray(mine[c0:c0 + Lc])).to(device)
                    rates = torch.empty((N, Lc), dtype=torch.float64, device=device)
                    y = torch.empty_like(rates); dy = torch.empty_like(rates); diag = torch.empty_like(rates)
                    means = torch.empty(Lc, dtype=torch.float64, device=device)
                    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
                    e0.record(stream)
                    _native.check(lib.mtg_sim_generate_dev(h, Lc, ctypes.byref(m0), p_dev.data_ptr(), None,
                                                           int(seed) * 1000003 + (i0 + c0), rates.data_ptr(), sp))
                    _native.check(lib.mtg_sim_add_noise_dev(h, Lc, rates.data_ptr(), int(seed) * 1000003 + 17 + (i0 + c0),
                                                            y.data_ptr(), dy.data_ptr(), sp))
                    _native.check(lib.mtg_prepare_data_dev(N, Lc, y.data_ptr(), dy.data_ptr(), diag.data_ptr(), means.data_ptr(), sp))
                    e1.record(stream)
                    # negative expected counts (rate * exposure + background < 0: Gaussian-PDF simulations of faint / highly
                    # variable sources) leave NaN in the lightcurve, hence in its mean; numpy raises ValueError("lam < 0")
                    # for the reference (noise_models.py:72).  Such simulations are flagged (status 3, T = NaN), counted and reported.
                    invalid = ~torch.isfinite(means)
                    # ---- null fits: start 0 = the generating parameters, others = 10 % Gaussian balls ------------
                    x0 = multistart_points(mine[c0:c0 + Lc], list(zip(lo0, hi0)), n_starts_null, rng)
                    B0 = Lc * n_starts_null
                    x0_dev = torch.from_numpy(np.ascontiguousarray(x0)).to(device)
                    ox = torch.empty((B0, nd0), dtype=torch.float64, device=device); of = torch.empty(B0, dtype=torch.float64, device=device)
                    oi = torch.empty(B0, dtype=torch.int32, device=device); os_ = torch.empty(B0, dtype=torch.int32, device=device)
                    _native.check(lib.mtg_lbfgs_dev(ctypes.byref(m0), N, Lc, n_starts_null, B0, t_dev.data_ptr(), y.data_ptr(), diag.data_ptr(),
                                                    means.data_ptr(), x0_dev.data_ptr(), ctypes.byref(opts), ox.data_ptr(), of.data_ptr(),
                                                    oi.data_ptr(), os_.data_ptr(), _native.FLAG_DS_MEAN, sp))
                    best0 = torch.empty(Lc, dtype=torch.float64, device=device); arg0 = torch.empty(Lc, dtype=torch.int32, device=device)
                    _native.check(lib.mtg_group_best_dev(Lc, n_starts_null, of.data_ptr(), best0.data_ptr(), arg0.data_ptr(), sp))
                    a0 = arg0.cpu().numpy().astype(np.int64)
                    x_null = ox.cpu().numpy()[a0]
                    ll_null = -best0.cpu().numpy()
                    if refine:
                        ll_null, x_null = mcmc_refine(m0, nd0, lo0, hi0, x_null, ll_null, Lc, y, diag, means, rng_ref, sp, 1, i0 + c0)
                    e2.record(stream)
                    # ---- alternative fits ---------------------------------------------------------------------------
                    x1 = alternative_starts(x_alt0, x_null, null_slots, lo1, hi1, n_starts_alt, rng).reshape(Lc * n_starts_alt, nd1)
                    B1 = Lc * n_starts_alt
                    x1_dev = torch.from_numpy(np.ascontiguousarray(x1)).to(device)
                    ax = torch.empty((B1, nd1), dtype=torch.float64, device=device); af = torch.empty(B1, dtype=torch.float64, device=device)
                    ai = torch.empty(B1, dtype=torch.int32, device=device); as_ = torch.empty(B1, dtype=torch.int32, device=device)
                    _native.check(lib.mtg_lbfgs_dev(ctypes.byref(m1), N, Lc, n_starts_alt, B1, t_dev.data_ptr(), y.data_ptr(), diag.data_ptr(),
                                                    means.data_ptr(), x1_dev.data_ptr(), ctypes.byref(opts), ax.data_ptr(), af.data_ptr(),
                                                    ai.data_ptr(), as_.data_ptr(), _native.FLAG_DS_MEAN, sp))
                    best1 = torch.empty(Lc, dtype=torch.float64, device=device); arg1 = torch.empty(Lc, dtype=torch.int32, device=device)
                    _native.check(lib.mtg_group_best_dev(Lc, n_starts_alt, af.data_ptr(), best1.data_ptr(), arg1.data_ptr(), sp))
                    a1 = arg1.cpu().numpy().astype(np.int64)
                    x_alt = ax.cpu().numpy()[a1]
                    ll_alt = -best1.cpu().numpy()
                    if refine:
                        ll_alt, x_alt = mcmc_refine(m1, nd1, lo1, hi1, x_alt, ll_alt, Lc, y, diag, means, rng_ref, sp, 2, i0 + c0)
                    e3.record(stream)
                    stream.synchronize()
                    local["simulate_ms"] += e0.elapsed_time(e1); local["fit_null_ms"] += e1.elapsed_time(e2); local["fit_alt_ms"] += e2.elapsed_time(e3)
                    sl = slice(c0, c0 + Lc)
                    inv = invalid.cpu().numpy()
                    out_l0[sl] = np.where(inv, np.nan, ll_null); out_l1[sl] = np.where(inv, np.nan, ll_alt)
                    out_p0[sl] = x_null; out_p1[sl] = x_alt
                    out_s0[sl] = np.where(inv, 3, os_.cpu().numpy()[a0]); out_s1[sl] = np.where(inv, 3, as_.cpu().numpy()[a1])
                    local["n_invalid"] += int(inv.sum())
                    if return_lightcurves:
                        kept[c0] = (y.cpu().numpy().T.copy(), dy.cpu().numpy().T.copy())
        finally:
            sim.close()
        return local

    n_slots = min(pipeline, len(starts))
    slots = [starts[k::n_slots] for k in range(n_slots)]
    if n_slots <= 1:
        results = [worker(slots[0])] if slots else []
    else:
        with concurrent.futures.ThreadPoolExecutor(max_workers=n_slots) as ex:
            results = list(ex.map(worker, slots))
    torch.cuda.synchronize(device)
    for r in results:
        for k in times:
            times[k] += r[k]
    out_T[:] = lrt_statistic(out_l0, out_l1)
    if times["n_invalid"]:
        import warnings
        warnings.warn("%d of %d simulated lightcurves had negative expected counts (the reference raises ValueError('lam < 0') "
                      "there); they are excluded from the LRT distribution (status 3, T = NaN)" % (times["n_invalid"], L))
    timings = dict(times, n_local=L, chunk=chunk, pipeline=n_slots,
                   note="per-stage device times are summed over concurrently running chunks")

    if dist is not None and world > 1:
        pack = np.empty((L, 5 + nd0 + nd1))
        pack[:, 0] = out_T; pack[:, 1] = out_l0; pack[:, 2] = out_l1; pack[:, 3] = out_s0; pack[:, 4] = out_s1
        pack[:, 5:5 + nd0] = out_p0; pack[:, 5 + nd0:] = out_p1
        allp = all_gather_rows(dist, pack, n_total, device)
        res = LRTResult(allp[:, 0], allp[:, 1], allp[:, 2], allp[:, 5:5 + nd0], allp[:, 5 + nd0:],
                        allp[:, 3].astype(np.int32), allp[:, 4].astype(np.int32), timings)
    else:
        res = LRTResult(out_T, out_l0, out_l1, out_p0, out_p1, out_s0, out_s1, timings)
    if return_lightcurves:
        ks = sorted(kept)
        res.lightcurves = (np.concatenate([kept[k][0] for k in ks]), np.concatenate([kept[k][1] for k in ks])) if ks else None
    return res
