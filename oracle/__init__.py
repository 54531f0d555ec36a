"""CPU oracle for the REED image hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32/fp64) functional restatement of the reference algorithm:

* ``sit_oracle``      <- /root/reference/image/models/sit.py   (+ the timm classes it imports)
* ``loss_oracle``     <- /root/reference/image/loss.py
* ``samplers_oracle`` <- /root/reference/image/samplers.py
* ``train_oracle``    <- /root/reference/image/train.py:84-105,363-385,396-412 (step glue, curriculum scalars)
* ``preprocess_oracle`` <- /root/reference/image/train.py:53-74 (raw-image preprocessing, tap-level bicubic)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``reed_b200``) never does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by ``oracle/make_golden.py``
(which imports the unmodified reference with ``oracle/timm_shim`` on ``sys.path``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks the oracle against those fixtures,
``tests/test_golden_reproducible.py`` that the fixtures regenerate bit-identically from the reference, and
``tests/test_oracle_vs_reference_live.py`` sweeps the oracle against the reference executed live.
"""
