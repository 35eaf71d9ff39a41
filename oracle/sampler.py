"""Oracle: emcee 3.1.4's ensemble sampler as ``GPModelling.derive_posteriors`` drives it.  TEST INFRASTRUCTURE ONLY.

emcee is an un-vendored dependency of the reference (``emcee==3.1.4``, ``/root/reference/pyproject.toml:13``) and is not
importable here; this module restates, from the published algorithm (Goodman & Weare 2010; Foreman-Mackey et al. 2013)
and emcee 3.1.4's documented behaviour, what the reference's call sites execute
(``/root/reference/mind_the_gaps/gpmodelling.py:247-264``: ``EnsembleSampler(walkers, ndim, self._log_probability,
pool=pool)`` then ``sampler.sample(p0, iterations=max_steps)`` with the default ``moves=StretchMove(a=2.0)``):

* ``EnsembleSampler.sample``: evaluate the log-probability of every start position once, then per iteration one
  ``RedBlueMove.propose``; the chain stores the ensemble after each iteration (``get_chain`` -> ``(steps, walkers, ndim)``).
* ``RedBlueMove.propose`` (``nsplits = 2``, ``randomize_split = True``): ``inds = arange(nwalkers) % 2`` is shuffled by
  the sampler's RandomState; for ``split`` in ``(0, 1)``: ``S1 = inds == split``; the walkers of ``S1`` (ascending index) are
  updated against the complementary set (ascending index) -- ``get_proposal`` -- all proposals of the half are evaluated in
  ONE call of the log-probability, then walker ``j`` (with factor ``f`` and new log-probability ``nlp``) accepts iff
  ``f + nlp - log_prob[j] > log(random.rand())``, one ``rand()`` per walker in ascending order.
* ``StretchMove.get_proposal(s, c, random)``: ``zz = ((a - 1) * random.rand(Ns) + 1) ** 2 / a``;
  ``factors = (ndim - 1) * log(zz)``; ``rint = random.randint(Nc, size=(Ns,))``; ``q = c[rint] - (c[rint] - s) * zz[:, None]``.

"parity unpinned" against the real package (not importable; no reference test pins a chain).  The random draws are
INPUTS here (``draws_like_emcee`` produces them from a ``numpy.random.RandomState`` in the order emcee consumes its
stream), so the device samplers K8 / K10 can be fed the same numbers through ``mtg_ensemble_injected_host`` and must
walk the same chain bit for bit (tests/test_ensemble_gpu.py).
"""
import numpy as np


def draws_like_emcee(random_state, nwalkers, nsteps):
    """The draws of ``nsteps`` iterations in the order emcee 3.1.4 takes them from its RandomState: per iteration
    ``shuffle(inds)``; per half ``rand(Ns)``, ``randint(Nc, size=Ns)``, then ``Ns`` single ``rand()`` for the accept tests.
    Returns ``split (nsteps, W) int32``, ``uz, uacc (nsteps, 2, H) float64``, ``rint (nsteps, 2, H) int32``, ``H = ceil(W/2)``
    (entries beyond ``Ns`` of the second half of an odd ensemble are unused and zero)."""
    W = int(nwalkers)
    H = (W + 1) // 2
    split = np.empty((nsteps, W), dtype=np.int32)
    uz = np.zeros((nsteps, 2, H)); uacc = np.zeros((nsteps, 2, H))
    rint = np.zeros((nsteps, 2, H), dtype=np.int32)
    for it in range(nsteps):
        inds = np.arange(W) % 2
        random_state.shuffle(inds)
        split[it] = inds
        for h in range(2):
            ns = int(np.sum(inds == h)); nc = W - ns
            uz[it, h, :ns] = random_state.rand(ns)
            rint[it, h, :ns] = random_state.randint(nc, size=(ns,))
            uacc[it, h, :ns] = [random_state.rand() for _ in range(ns)]
    return split, uz, rint, uacc


def stretch_move_run(log_prob_fn, p0, nsteps, split, uz, rint, uacc, a=2.0):
    """``nsteps`` iterations of emcee's stretch move from ``p0 (W, ndim)`` on the given draws.

    ``log_prob_fn``: ``(B, ndim) -> (B,)`` (``-inf`` allowed, NaN raises as in emcee).
    Returns ``chain (nsteps, W, ndim)``, ``chain_lp (nsteps, W)``, ``n_accept (W,)``."""
    pos = np.array(p0, dtype=np.float64)
    W, ndim = pos.shape
    lp = np.asarray(log_prob_fn(pos), dtype=np.float64).copy()
    if np.any(np.isnan(lp)):
        raise ValueError("The initial log-probability is NaN")
    chain = np.empty((nsteps, W, ndim)); chain_lp = np.empty((nsteps, W))
    n_accept = np.zeros(W, dtype=np.int64)
    all_inds = np.arange(W)
    for it in range(nsteps):
        inds = np.asarray(split[it])
        for h in range(2):
            S1 = inds == h
            s = pos[S1]
            c = pos[~S1]                                    # sets[:split] + sets[split + 1:] for nsplits = 2
            ns, nc = len(s), len(c)
            zz = ((a - 1.0) * uz[it, h, :ns] + 1) ** 2.0 / a
            factors = (ndim - 1.0) * np.log(zz)
            ri = np.asarray(rint[it, h, :ns])
            q = c[ri] - (c[ri] - s) * zz[:, None]
            new_lp = np.asarray(log_prob_fn(q), dtype=np.float64)
            if np.any(np.isnan(new_lp)):
                raise ValueError("Probability function returned NaN")
            for i, (j, f, nlp) in enumerate(zip(all_inds[S1], factors, new_lp)):
                lnpdiff = f + nlp - lp[j]
                with np.errstate(divide="ignore"):
                    if lnpdiff > np.log(uacc[it, h, i]):
                        pos[j] = q[i]; lp[j] = nlp; n_accept[j] += 1
        chain[it] = pos; chain_lp[it] = lp
    return chain, chain_lp, n_accept
