"""
femvf_b200 -- B200-native hot path of ``femvf.forward.integrate`` behind the femvf model API.

Module layout mirrors the reference package (``/root/reference/src/femvf``): ``forward``,
``statefile``, ``load``, ``static``, ``models.transient``, ``models.fsi``, ``residuals.solid``,
``residuals.fluid``, ``equations.newmark``, ``solverconst``, ``meshutils``.  Compute goes
through the C ABI of ``lib/libvffem_b200.so`` (``include/vffem_b200.h``); see DESIGN.md.
"""

__all__ = ['blockvec', 'mesh', 'meshgen', 'meshutils', 'load', 'forward', 'statefile', 'static',
           'solverconst']
