"""TEST INFRASTRUCTURE: drives the host-compiled frame functions (tests/host_harness/emu.cpp) through the same
bookkeeping code the product uses (glamr_b200.problem), with the oracle's SMPL supplying the joints."""
import ctypes

import numpy as np
import torch

import host_harness as hh
from glamr_b200 import lib as L
from glamr_b200 import problem as PB
from oracle import rotations as rt


def _fp(t):
    return ctypes.c_void_p(t.data_ptr())


class EmuRunner:
    def __init__(self, oracle_model, data):
        self.model, self.data = oracle_model, data
        self.flags = {k: getattr(oracle_model, k) for k in
                      ['flag_fixed_cam', 'flag_opt_cam', 'flag_opt_cam_from_person_pose', 'flag_cam_inv_trans_res_all',
                       'flag_opt_vis_local_rot', 'cam_fix_frames']}
        self.layout = PB.make_layout(data, self.flags)
        self.theta = torch.zeros(self.layout.n_params)
        PB.bind_variables(data, self.layout, self.theta)
        self.comp = PB.StageCompiler(data, self.layout, self.flags, 'cpu', rt.aa_to_rot6d, aa_to_quat=rt.aa_to_quat)
        self.lib = hh.lib()
        self.h = None
        self.reduce = torch.zeros(self.layout.n_params + L.NUM_TERMS)

    def set_stage(self, opt_variables, loss_cfg, stage, reset_adam=True, **kw):
        PB.begin_stage_variables(self.data, self.layout, self.theta, self.flags, opt_variables)
        self.pb = self.comp.compile(self.theta, opt_variables, loss_cfg, stage, **kw)
        if self.h is None:
            self.h = ctypes.c_void_p()
            assert self.lib.glamr_host_emu_create(ctypes.byref(self.h), ctypes.byref(self.pb)) == 0
        else:
            assert self.lib.glamr_host_emu_set_problem(self.h, ctypes.byref(self.pb), int(reset_adam)) == 0

    def buffer(self, what):
        p, n = ctypes.POINTER(ctypes.c_float)(), ctypes.c_size_t()
        assert self.lib.glamr_host_emu_buffer(self.h, what, ctypes.byref(p), ctypes.byref(n)) == 0
        return torch.from_numpy(np.ctypeslib.as_array(p, shape=(n.value,)))

    def backward(self):
        lib, P, T = self.lib, self.comp.P, self.comp.T
        assert lib.glamr_host_emu_forward_pose(self.h, _fp(self.theta)) == 0
        ow = self.buffer(L.R_ORIENT_WORLD).view(P * T, 3).clone()
        tw = self.buffer(L.R_TRANS_WORLD).view(P * T, 3).clone()
        scale = None if self.comp.scale_all is None else self.comp.scale_all.reshape(-1)
        joints, _ = self.model.smpl(ow, self.comp.pose_all.reshape(P * T, 69), self.comp.beta_all.reshape(P * T, 10),
                                    root_trans=tw, root_scale=scale)
        self.buffer(L.R_JOINTS_WORLD).copy_(joints.reshape(-1))
        assert lib.glamr_host_emu_backward(self.h, _fp(self.theta), _fp(self.reduce)) == 0
        terms = torch.zeros(L.NUM_TERMS + 1)
        assert lib.glamr_host_emu_losses(self.h, _fp(self.reduce), _fp(terms)) == 0
        return self.reduce[:self.layout.n_params], terms

    def step(self, lr):
        assert self.lib.glamr_host_emu_adam(self.h, _fp(self.theta), _fp(self.reduce), ctypes.c_double(lr)) == 0
