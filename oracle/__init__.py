"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / torch-CPU restatement of the reference's algorithm
for the two in-scope paths (distill.py inner loop, epoch.py retrieval scoring).
It exists to check the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing
under ``multimodal_dataset_distillation_b200/`` imports it, and the product path
raises if the CUDA library is missing -- there is no CPU fallback.

Contents: ``retrieval_ref.py`` (numpy), ``distill_ref.py`` (torch CPU, fp64 capable) and ``c/itm_eval_ref.c``, a plain-C
restatement of itm_eval's integer part (ranks with the index tie-break, recall@1/5/10) built by ``c/Makefile`` into
``_build/libitm_port.so`` (the builder's own port, not the reference compiled) -- an independent cross-check of the numpy one (``retrieval_ref.itm_eval_c``).

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this
path ("parity unpinned" upstream).  The oracle is therefore
pinned against outputs of the reference itself, run in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference/epoch.py``,
``epoch_original.py`` and ``reparam_module.py`` (the importable parts), runs them on
seeded inputs and commits the results under ``tests/golden/``.  ``distill.py`` and
``networks.py`` cannot be imported (clip/timm/kornia missing, BERT download at
import), so the inner loop is restated from distill.py:509-613 and driven through
the real ``ReparamModule`` + torch double-backward when the goldens are made.
"""
