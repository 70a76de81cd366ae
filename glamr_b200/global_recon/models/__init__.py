"""Same registry as the reference's ``global_recon/models/__init__.py:4-6``: ``cfg.grecon_model_name`` selects the class."""
from glamr_b200.recon import GlobalReconOptimizer

model_dict = {
    'global_recon_model': GlobalReconOptimizer
}
