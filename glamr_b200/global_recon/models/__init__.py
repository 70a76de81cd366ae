"""Optimiser registry of the drop-in package.

GLAMR's entry points look the optimiser class up by ``cfg.grecon_model_name`` (run_demo.py:59, run_dataset.py:64 of the
reference); the only name its configs use is ``global_recon_model``, which resolves to the CUDA-backed class here.
"""


def _build_registry():
    from glamr_b200.recon import GlobalReconOptimizer as cuda_optimizer
    registry = dict()
    registry['global_recon_model'] = cuda_optimizer
    return registry


model_dict = _build_registry()
