"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / torch-CPU restatement of the reference's algorithm
for the two in-scope paths (distill.py inner loop, epoch.py retrieval scoring).
It exists to check the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing
under ``multimodal_dataset_distillation_b200/`` imports it, and the product path
raises if the CUDA library is missing -- there is no CPU fallback.

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
