"""CPU oracle for the SuperDiff sampling path.  TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a CPU (PyTorch / NumPy, fp64 and fp32) restatement of the
reference's sampling math.  Every function cites the reference file:line it
follows (paths relative to the reference checkout, ``*.ipynb:N`` = line N of
the raw notebook JSON).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``super_diffusion_b200/``
imports it; the product path raises when the CUDA extension is missing.

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer
fixtures for this path, and its own implementation (JAX / Flax /
diffusers) cannot be imported in this image.  The oracle is
pinned only by (a) line-by-line restatement of the cited sources, (b) the
algebraic identities of SURVEY.md Appendix A checked in fp64
(tests/test_oracle_identities.py) and (c) golden vectors generated *by this
oracle* (tests/golden/, generator script committed).  Third-party arithmetic
that is restated from documented semantics: flax==0.9.0 ``nn.Conv`` (SAME
padding), ``nn.GroupNorm`` (32 groups, eps 1e-6), ``nn.Dense``, ``nn.Embed``,
``jax.nn.softmax``, ``jax.image.resize('nearest')`` (cifar/requirements.txt:24,39)
and diffusers' ``EulerDiscreteScheduler`` sigma table (version unpinned).
"""
