"""Mirror of the reference's ``pdm`` package surface for the hot path:
``pdm.models`` (pruned/gated U-Net constructors, arch-vector plumbing), ``pdm.losses``, ``pdm.utils.compute_snr`` and
the step/optimizer pieces of ``pdm.training``.  Everything numerical dispatches to ``libb200pdm.so``."""
__version__ = "2.2.0+b200"
