"""Import shim for the (absent, unpinned) third-party `smplx` package: an SMPL module that loads the synthetic
npz assets and restates upstream smplx.SMPL.forward (lbs -> 21 vertex picks -> optional transl)."""
import numpy as np
import torch
import torch.nn as nn
from collections import namedtuple
from .lbs import lbs

SMPLOutput = namedtuple('SMPLOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])
_EXTRA_V = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
            2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]


class SMPL(nn.Module):
    def __init__(self, model_path, *a, **kw):
        super().__init__()
        d = np.load(model_path + '/SMPL_SYNTH.npz')
        for k in ['v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights']:
            self.register_buffer(k, torch.tensor(d[k], dtype=torch.float32))
        self.register_buffer('parents', torch.tensor(d['parents'], dtype=torch.long))
        self.faces = d['faces']

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_full_pose=False,
                pose2rot=True, **kw):
        if global_orient is None:
            global_orient = torch.zeros_like(body_pose[:, :3])
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        B = max(betas.shape[0], full_pose.shape[0])
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        v, j = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
                   self.parents, self.lbs_weights)
        j = torch.cat([j, v[:, _EXTRA_V]], dim=1)
        if transl is not None:
            j = j + transl[:, None]
            v = v + transl[:, None]
        return SMPLOutput(v, j, full_pose, betas, global_orient, body_pose)
