"""CPU oracle for the BayesSimIG hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference
(NVlabs/bayes-sim-ig) for the path this repository accelerates:

    trajectory summarizers -> MDNN / MDRFF mixture-density training ->
    mixture-of-Gaussians posterior evaluation and sampling.

Nothing under ``bayes_sim_ig_b200/`` (the product) imports this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
CPU timing baseline -- never as the thing that is shipped.

Parity pinning
--------------
* summarizers (start / waypts / corr / corrdiff), MDNN / MDRFF forward, loss,
  gradients, Adam, ``predict_MoGs``, ``Gaussian(L=)``, ``MoG.gen`` and
  ``MoG.eval``: **pinned** against the reference's own torch / numpy code,
  imported live from ``/root/reference`` by ``tests/golden/make_golden.py``;
  the resulting vectors are committed under ``tests/golden/*.npz`` and
  ``tests/test_oracle_golden.py`` checks the oracle against them.
* ``ParamsGenerator.sample`` per environment (the posterior consumer,
  sim/params_generator.py:115-118): **pinned**, ``pdf_np.params_samples_per_env`` is
  bit-exact against 41 sequential calls of the live reference (``*.envs.*`` in
  ``tests/golden/pdf.npz``).
* path signature (``summary_signatory``): **parity unpinned**.  The arithmetic
  lives in the third-party ``signatory`` package (patrick-kidger/signatory,
  un-pinned by the reference: README.md:345-350 installs master), which is
  absent here and cannot be installed.  The oracle restates the published
  definition (Chen product of per-increment tensor exponentials) in float64
  and is validated by algebraic identities and by an independent
  iterated-sum formulation (``oracle/signature_np.py``).
* ``ghalton`` (quasi-random RFF frequencies): **parity unpinned**, absent.
  Frequencies are constructor-time constants; parity tests copy them.
"""
