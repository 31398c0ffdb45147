"""CPU oracle for the PEViT fine-tuning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pevit_b200/`` may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker (or
as the timed CPU baseline), never as part of the product path.

The oracle restates, in plain PyTorch CPU ops, the algorithm of the reference
(eric-ai-lab/PEViT @ be6fb43, ``vision_benchmark/evaluation/{model,lora_model,
adapter_model,compacter_model}.py``).  It follows the reference's *own* way of
computing things (materialised Kronecker sums, dense DxD delta matmul, the raw
``reshape`` scramble, the duplicated MLP evaluation), not the factorised math
the CUDA kernels use, so that agreement between the two is meaningful.

Parity pin: the reference has no tests / golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` (``oracle/ref_import.py``), runs
them on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks the oracle against those fixtures and,
when ``/root/reference`` is mounted, against the live reference as well.
"""
