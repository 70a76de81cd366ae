"""Which chain bounds the graph-replayed iteration?  Experiment build only: times the warm (back-to-back) and L2-flushed iteration
with single kernels left out (GLAMR_EXP_SKIP bits: 1 blend, 2 skinning, 4 residuals, 8 backward).  Results of a skipping run are
meaningless as numbers of the optimisation; only the timing differences are read.

    GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=<mask> python tools/iter_skip_exp.py
"""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glamr_b200 import lib as L
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.synthetic import make_in_dict, make_smpl_assets, SyntheticPrior
P, T = int(os.environ.get('P', 1)), int(os.environ.get('T', 300))
cfgid = os.environ.get('CFG', 'glamr_dynamic')
a = make_smpl_assets(0); dev = torch.device('cuda:0')
cfg = Config(cfgid); in_dict = make_in_dict(a, P, T)
m = GlobalReconOptimizer(cfg, dev, None, smpl=SMPL(a, device=dev), mt_model=SyntheticPrior(0, dev))
data = m.init_data(copy.deepcopy(in_dict))
stage, specs = list(cfg.opt_stage_specs.items())[-1]
m._cur_vars, m._cur_stage, m._loss_cfg = specs['opt_variables'], stage, specs['loss_cfg']
m._set_stage(data, specs['opt_variables'], specs['loss_cfg'], stage, reset_adam=True, begin=True)
hist = torch.zeros((4000, L.NUM_TERMS + 1), device=dev)
lib = m._lib
def it(n, graph):
    L.check(lib.glamr_opt_iterate(m._opt, L.ptr(m._theta), L.ptr(m._reduce), float(specs['opt_lr']), L.ptr(hist), L.NUM_TERMS + 1, n, graph, L.stream_ptr()), 'iterate')
it(5, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 200
e0.record(); it(K, 1); e1.record(); torch.cuda.synchronize()
warm = e0.elapsed_time(e1) / K * 1e3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cold = 0.0
for _ in range(50):
    flush.fill_(1)
    e0.record(); it(1, 1); e1.record(); torch.cuda.synchronize()
    cold += e0.elapsed_time(e1) * 1e3 / 50
print(f'P={P} T={T} {cfgid}:{stage} skip={os.environ.get("GLAMR_EXP_SKIP", "0")}: warm {warm:.1f} us / iteration, L2-flushed {cold:.1f} us')
