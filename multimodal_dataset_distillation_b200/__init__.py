"""B200-native hot path of vision-language trajectory-matching dataset distillation.

Public surface (mirrors the reference's names for the accelerated path):
  reparam_module.ReparamModule      flat-parameter functional wrapper (reparam_module.py)
  epoch.itm_eval / epoch_test / evaluate_synset / epoch   (epoch.py, epoch_original.py)
  distill.main / distill.build_parser / distill.DistillEngine   (distill.py)
  ops.*                             tensor-level entry points of the C ABI (include/vldd_b200.h)
"""
__version__ = "0.1.0"
