"""CPU oracle for the DiffSHEG sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as the thing shipped.  The product path
(``diffsheg_b200``) fails loudly when its CUDA library is missing instead of
falling back to anything here.

What it is: a functional (state_dict in, tensors out) torch-CPU restatement of

* ``models/transformer.py``  UniDiffuser.forward  (tr:728-770) and everything
  below it (MotionTransformer tr:496-587, the linear-attention layer
  tr:300-346, StylizationBlock tr:86-97, FFN tr:178-181, timestep_embedding
  tr:42-59, PeriodicPositionalEncoding tr:19-38), in ``oracle/denoiser.py``;
* ``models/gaussian_diffusion.py`` coefficient tables gd:351-387, ``_undo``
  gd:467-473, ``p_mean_variance`` gd:499-612, ``p_sample`` gd:684-774,
  ``ddim_sample`` gd:976-1066 and the four sample loops gd:776-974,
  gd:1106-1278; ``models/respace.py`` rs:7-124; ``models/scheduler.py``
  sch:47-62,150-209, in ``oracle/diffusion.py``.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference ITSELF, imported from /root/reference in
the build container: ``tests/golden/make_golden.py`` runs the real
``UniDiffuser`` / ``SpacedDiffusion`` on seeded synthetic weights and inputs and
commits the input/output vectors under ``tests/golden/``;
``tests/test_oracle.py`` checks the oracle against those fixtures everywhere and
against the live reference whenever /root/reference is present.
"""
