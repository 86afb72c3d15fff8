"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's NeRF hot path (``ngp_oracle.c`` + numpy wrappers in
``oracle.py``), the recipe that builds the reference's own CUDA extensions as the strongest
checker (``build_ref.py`` -> ``oracle/_ref/``), and a CPU port of the reference's pure-PyTorch
renderer used as the reported CPU baseline (``torch_renderer.py``).

Nothing under ``single-stable-dreamfusion_b200/`` imports this package.  Allowed callers:
``tests/``, ``__graft_entry__.smoke()/build()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.
"""
