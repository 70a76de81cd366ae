"""Learned motion / trajectory prior on the CUDA library -- drop-in for the inference surface of the reference's
``MotionTrajJointModel`` (motion_infiller/models/motion_traj_joint_model.py:17-145): ``inference(batch, sample_num)``,
``get_motion_latent``, ``get_traj_latent``.

The networks run in ``glamr_infiller_forward`` (the autoregressive sweep of 50-frame windows of
MotionInfillerVAE.inference_multi_step, motion_infiller_vae.py:618-632, batched over ALL sequences instead of the
reference's batch of one) and ``glamr_trajpred_forward`` (glamr_b200/csrc/nets_kernels.cu); this module runs SMPL FK for the joint-position features and reshapes the outputs into the reference's
dict layout.  Weights come from the reference's Lightning checkpoints (state_dict names are kept) or from an explicit
state dict; there is no CPU fallback.
"""
import ctypes
import glob
import os

import numpy as np
import torch

from . import lib as L

PAST, CUR, FUT, NZ = 10, 30, 10, 128
PRIOR_GRAPH_DEFAULT = '0'
WINDOW = PAST + CUR + FUT


def _declare(lib):
    if getattr(lib, '_nets_declared', False):
        return
    lib.glamr_infiller_workspace_floats.restype = ctypes.c_size_t
    lib.glamr_infiller_sequence_workspace_floats.restype = ctypes.c_size_t
    lib.glamr_infiller_forward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                                                                     ctypes.c_void_p]
    lib.glamr_trajpred_workspace_floats.restype = ctypes.c_size_t
    lib.glamr_net_set_tensor.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.glamr_infiller_window_forward.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p,
                                                                                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.glamr_trajpred_forward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + \
        [ctypes.c_void_p] * 6 + [ctypes.c_size_t, ctypes.c_void_p]
    lib._nets_declared = True


class _GraphCache:
    """Replays a fixed launch sequence as ONE CUDA graph.  The prior networks are launch-latency bound (about 70 small kernels
    per 50-frame window, 10 windows for a 300-frame track): `run(key, fn, inputs)` copies `inputs` into static device
    buffers, replays the graph captured for `key` (same shapes -> same launches) and returns clones of the static outputs.
    The first call for a key runs `fn` eagerly (lazy initialisation must happen outside capture) and then captures it; if
    capture is refused the key stays on the eager path."""

    def __init__(self, enabled=None, max_entries=8):
        if enabled is None:          # GLAMR_PRIOR_GRAPH=1|0; the default flips to on once verified on a B200
            enabled = os.environ.get('GLAMR_PRIOR_GRAPH', PRIOR_GRAPH_DEFAULT) == '1'
        self.enabled, self.max_entries, self.entries = enabled, max_entries, {}

    def run(self, key, fn, inputs):
        if not self.enabled:
            return fn(*inputs)
        ent = self.entries.get(key)
        if ent is None:
            if len(self.entries) >= self.max_entries:
                self.entries.pop(next(iter(self.entries)))
            static_in = [None if x is None else x.clone() for x in inputs]
            out = fn(*static_in)                                   # eager warm-up; also the result of this first call
            ent = {'in': static_in, 'graph': None, 'out': None}
            self.entries[key] = ent
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ent['out'] = fn(*static_in)
                ent['graph'] = g
            except Exception:
                ent['graph'], ent['out'] = None, None
                torch.cuda.synchronize()
            return out
        if ent['graph'] is None:
            return fn(*inputs)
        for dst, src in zip(ent['in'], inputs):
            if dst is not None:
                dst.copy_(src)
        ent['graph'].replay()
        return tuple(o.clone() for o in ent['out']) if isinstance(ent['out'], tuple) else ent['out'].clone()


class _Net:
    """opaque glamr_net_t with the parameters of one network"""

    def __init__(self, state, device):
        self.device = L.require_cuda(device)
        self.lib = L.load()
        _declare(self.lib)
        self.h = ctypes.c_void_p()
        L.check(self.lib.glamr_net_create(ctypes.byref(self.h)), 'glamr_net_create')
        with torch.cuda.device(self.device):
            for name, val in state.items():
                arr = np.ascontiguousarray(val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else np.asarray(val), dtype=np.float32)
                if arr.size == 0 or arr.dtype != np.float32:
                    continue
                L.check(self.lib.glamr_net_set_tensor(self.h, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), arr.size), f'set_tensor {name}')
        self._ws = None

    def workspace(self, floats):
        if self._ws is None or self._ws.numel() < floats:
            if self._ws is not None:                 # captured graphs may hold the old address: retire, do not free
                self._retired = getattr(self, '_retired', []) + [self._ws]
            self._ws = torch.empty(int(floats), dtype=torch.float32, device=self.device)
        return self._ws

    def __del__(self):
        try:
            self.lib.glamr_net_destroy(self.h)
        except Exception:
            pass


class MotionInfillerVAE:
    """inference surface of motion_infiller/models/motion_infiller_vae.py:440-667 (pose_rep 'body', axis-angle)"""
    model_type = 'angle'

    def __init__(self, state, device):
        self.net = _Net(state, device)
        self.device = self.net.device
        self.nz, self.past_nframe, self.cur_nframe, self.fut_nframe = NZ, PAST, CUR, FUT
        self.graphs = _GraphCache()

    def get_latent(self, seq_len):
        return torch.randn((int(np.ceil((seq_len - PAST) / CUR)), NZ))

    def inference(self, batch, sample_num=1, recon=False, multi_step=True):
        if recon or not multi_step:
            raise NotImplementedError('only the multi-step sampling path (recon=False) is implemented on CUDA')
        dev = self.device
        pose_in = batch['in_body_pose'].to(dev, torch.float32).contiguous()
        frame_mask = batch['frame_mask'].to(dev, torch.float32).contiguous()
        latent = batch.get('in_motion_latent')
        latent = None if latent is None else latent.to(dev, torch.float32).contiguous()
        B0, T = pose_in.shape[:2]
        self.net.workspace(self.net.lib.glamr_infiller_sequence_workspace_floats(B0 * sample_num))        # sized before any capture
        key = ('infill', B0, T, sample_num, None if latent is None else tuple(latent.shape))
        with torch.cuda.device(dev):
            body = self.graphs.run(key, lambda a, b, c: self._windows(a, b, c, sample_num), (pose_in, frame_mask, latent))
        data = dict(batch)
        data['infer_out_body_pose'] = body
        data['infer_out_pose'] = torch.cat([torch.zeros_like(body[..., :3]), body], dim=-1)
        data['batch_size'], data['seq_len'] = B0, T
        return data

    def _windows(self, pose_in, frame_mask, latent, sample_num):
        """motion_infiller_vae.py:618-632: autoregressive 50-frame windows, stride 30 -- one library call for the whole sweep
        (device tensors in, [B0, S, T, 69] out; no host sync)"""
        dev = self.device
        B0, T = pose_in.shape[:2]
        B = B0 * sample_num
        pose = pose_in.repeat_interleave(sample_num, dim=0).transpose(0, 1).contiguous().clone()       # [T,B,69], overwritten in place
        key_pad_all = (~(frame_mask == 1)).repeat_interleave(sample_num, dim=0).to(torch.uint8).contiguous()      # [B,T]
        nwin = int(np.ceil((T - PAST) / CUR))
        if latent is not None and latent.dim() == 3:                 # [B0, windows, nz]: one latent per sequence and window
            eps, rows = latent[:, :nwin].repeat_interleave(sample_num, dim=0).transpose(0, 1).contiguous(), B
        elif latent is not None:                                     # [windows, nz]: shared by the batch
            eps, rows = latent[:nwin].contiguous(), 1
        else:
            eps, rows = torch.randn((nwin, B, NZ), device=dev), B
        if eps.shape[0] < nwin:
            raise ValueError(f'{nwin} windows need {nwin} latents, got {eps.shape[0]}')
        lib = self.net.lib
        ws = self.net.workspace(lib.glamr_infiller_sequence_workspace_floats(B))
        L.check(lib.glamr_infiller_forward(self.net.h, T, B, pose.data_ptr(), key_pad_all.data_ptr(), eps.data_ptr(), rows, ws.data_ptr(), ws.numel(),
                                           torch.cuda.current_stream().cuda_stream), 'glamr_infiller_forward')
        return pose.transpose(0, 1).reshape(B0, sample_num, T, 69).contiguous()


class TrajPredVAE:
    """inference surface of traj_pred/models/traj_pred_vae.py:341-548 (6d local orientation, joint-position input)"""
    model_type = 'joint'
    in_joint_pos_only = False

    def __init__(self, state, device, smpl):
        self.net = _Net(state, device)
        self.device, self.smpl, self.nz = self.net.device, smpl, NZ
        self.graphs = _GraphCache()

    def get_latent(self, seq_len):
        return torch.zeros((1, NZ))

    def get_joint_pos(self, body_pose):
        """:384-394  23 FK joints (root removed), zero orientation, rest joints from v_template"""
        flat = body_pose.reshape(-1, 69).to(self.device, torch.float32).contiguous()
        z3 = torch.zeros((flat.shape[0], 3), device=self.device)
        joints = self.smpl.get_joints(global_orient=z3, body_pose=flat, root_trans=z3)
        return joints[:, 1:, :].reshape(body_pose.shape[:-1] + (69,))

    def inference(self, batch, sample_num=1, recon=False, recon_only=False, multi_step=False):
        if recon or recon_only or multi_step or sample_num != 1:
            raise NotImplementedError('only single-pass sampling (multi_step_trajpred=false, sample_num=1) is implemented on CUDA')
        dev = self.device
        body = batch['in_body_pose'].to(dev, torch.float32).contiguous()           # [B,T,69]
        B, T = body.shape[:2]
        dv = lambda k: batch[k].to(dev, torch.float32).contiguous() if k in batch and batch[k] is not None else None
        latent, ixy, ih = dv('in_traj_latent'), dv('init_xy'), dv('init_heading')
        self.net.workspace(self.net.lib.glamr_trajpred_workspace_floats(T, B))
        self.smpl._workspace(B * T, fk_only=True)
        key = ('traj', B, T, None if latent is None else tuple(latent.shape), ixy is not None, ih is not None)
        with torch.cuda.device(dev):
            local, trans, orient = self.graphs.run(key, self._forward, (body, latent, ixy, ih))
        out = {'infer_out_local_traj_tp': local.view(T, B, 1, 11), 'infer_out_trans_tp': trans.view(T, B, 1, 3),
               'infer_out_orient_tp': orient.view(T, B, 1, 3)}
        out['infer_out_orient'] = out['infer_out_orient_tp'].permute(1, 2, 0, 3).contiguous()
        out['infer_out_trans'] = out['infer_out_trans_tp'].permute(1, 2, 0, 3).contiguous()
        out['infer_out_pose'] = torch.cat([out['infer_out_orient'], body.unsqueeze(1)], dim=-1)
        return out

    def _forward(self, body, latent, ixy, ih):
        """joint-position features (SMPL FK) + the network, device tensors only"""
        dev = self.device
        B, T = body.shape[:2]
        jp = self.get_joint_pos(body).transpose(0, 1).contiguous()                # [T,B,69]
        lib = self.net.lib
        ws = self.net.workspace(lib.glamr_trajpred_workspace_floats(T, B))
        local = torch.empty((T, B, 11), dtype=torch.float32, device=dev)
        trans = torch.empty((T, B, 3), dtype=torch.float32, device=dev)
        orient = torch.empty((T, B, 3), dtype=torch.float32, device=dev)
        if latent is not None:
            eps, rows = latent, (1 if latent.shape[0] == 1 else B)
        else:
            eps, rows = torch.randn((B, NZ), device=dev), B
        L.check(lib.glamr_trajpred_forward(self.net.h, T, B, jp.data_ptr(), eps.data_ptr(), rows, None if ixy is None else ixy.data_ptr(),
                                           None if ih is None else ih.data_ptr(), local.data_ptr(), trans.data_ptr(), orient.data_ptr(),
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), 'glamr_trajpred_forward')
        return local, trans, orient


def load_lightning_state_dict(path):
    """state_dict of a PyTorch-Lightning ``.ckpt`` (lib/utils/tools.py:94-104 -> ``load_from_checkpoint``) as plain float32
    numpy arrays.  Lightning checkpoints pickle hyper-parameter objects next to the tensors; tensors-only loading
    (``weights_only=True``) is tried first so that nothing but tensors is unpickled, the permissive path is the fall-back
    for files that need it."""
    try:
        ck = torch.load(path, map_location='cpu', weights_only=True)
    except Exception:
        ck = torch.load(path, map_location='cpu', weights_only=False)
    sd = ck.get('state_dict', ck)
    return {k: (v.detach().float().numpy() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}


def _find_checkpoint(cfg_dir, cp='best', version=None):
    """lib/utils/tools.py:41-45 (find_last_version) + :94-104 (get_checkpoint_path): 'best' = the LAST of the sorted
    *best*.ckpt names, 'last' = last.ckpt, an integer = model-epoch=NNNN.ckpt"""
    if version is None:
        numbers = sorted(int(os.path.basename(p)[len('version_'):]) for p in glob.glob(os.path.join(cfg_dir, 'version_*')))
        if not numbers:
            raise FileNotFoundError(f'no checkpoint versions under {cfg_dir}')
        version = numbers[-1]
    ck_dir = os.path.join(cfg_dir, f'version_{version}', 'checkpoints')
    if cp == 'last':
        return os.path.join(ck_dir, 'last.ckpt')
    if cp == 'best':
        files = sorted(glob.glob(os.path.join(ck_dir, '*best*.ckpt')))
        if not files:
            raise FileNotFoundError(f'no *best*.ckpt under {ck_dir}')
        return files[-1]
    return os.path.join(ck_dir, f'model-epoch={int(cp):04d}.ckpt')


# motion_infiller/cfg_infer/joint_motion_traj_demo.yml (the only joint config the reference ships)
_JOINT_DEMO_YML = {
    'results_root_dir': 'results/motion_filler_infer', 'seed': 1,
    'model_specs': {'mfiller_cfg': 'motion_infiller_demo', 'mfiller_cp': 'best', 'trajpred_cfg': 'traj_pred_demo', 'trajpred_cp': 'best'},
    'amass_dir': 'datasets/amass_processed/v1', 'seq_len': 300, 'seq_sampling_method': 'length',
    'data_mask_methods': {'drop_frames': {'preserve_first_n': 10, 'min_drop_len': 5, 'max_drop_len': 200}},
    'num_motion_samp': 3, 'multi_step_mfiller': True, 'multi_step_trajpred': False,
}
_RESULTS_ROOT = {'motion_infiller': 'results/motion_filler', 'traj_pred': 'results/traj_pred'}   # results_root_dir of the two shipped network configs


class MTConfig:
    """motion_infiller/utils/config_motion_traj.py:7-45: the joint model's YAML (cwd-relative glob like the reference; the
    shipped joint_motion_traj_demo.yml is built in) and the checkpoint directories of the two networks it names
    (motion_infiller/utils/config.py:16-26, traj_pred/utils/config.py:16-26)."""

    def __init__(self, cfg_id):
        import yaml
        self.id = cfg_id
        files = glob.glob(f'motion_infiller/cfg_infer/**/{cfg_id}.yml', recursive=True)
        if len(files) == 1:
            self.yml_dict = yaml.safe_load(open(files[0]))
        elif cfg_id == 'joint_motion_traj_demo':
            self.yml_dict = {k: (dict(v) if isinstance(v, dict) else v) for k, v in _JOINT_DEMO_YML.items()}
        else:
            raise FileNotFoundError(f'motion_infiller/cfg_infer/**/{cfg_id}.yml not found')
        y = self.yml_dict
        self.model_specs = y.get('model_specs', {})
        self.seed = y.get('seed', 1)
        self.multi_step_mfiller = y.get('multi_step_mfiller', True)
        self.multi_step_trajpred = y.get('multi_step_trajpred', True)

    @staticmethod
    def network_cfg_dir(package, net_cfg_id):
        import yaml
        files = glob.glob(f'{package}/cfg/**/{net_cfg_id}.yml', recursive=True)
        root = _RESULTS_ROOT[package]
        if len(files) == 1:
            root = os.path.expanduser(yaml.safe_load(open(files[0])).get('results_root_dir', root))
        return f'{root}/{net_cfg_id}'


class MotionTrajJointModel:
    supports_person_batch = True     # inference() accepts [B, T, 69] with B > 1 (GlobalReconOptimizer.infer_motion_traj_all)

    def __init__(self, cfg=None, device=torch.device('cuda'), log=None, smpl=None, states=None):
        """cfg: config id / object of the joint model (only its checkpoint locations are used).  states: optional
        (infiller_state_dict, trajpred_state_dict); otherwise the reference's checkpoint files are loaded."""
        self.device, self.log = L.require_cuda(device), log
        if isinstance(cfg, str):
            cfg = MTConfig(cfg)
        self.cfg = cfg
        self.multi_step_mfiller = getattr(cfg, 'multi_step_mfiller', True)
        self.multi_step_trajpred = getattr(cfg, 'multi_step_trajpred', False)
        if self.multi_step_trajpred:
            raise NotImplementedError('multi_step_trajpred: chunked trajectory prediction is not implemented on CUDA (the shipped joint config disables it)')
        if smpl is None:
            from .smpl import SMPL
            smpl = SMPL(device=self.device)
        self.smpl = smpl
        if states is None:
            specs = getattr(cfg, 'model_specs', None) or _JOINT_DEMO_YML['model_specs']
            states = []
            for package, key in [('motion_infiller', 'mfiller'), ('traj_pred', 'trajpred')]:
                path = _find_checkpoint(MTConfig.network_cfg_dir(package, specs[f'{key}_cfg']), specs.get(f'{key}_cp', 'best'),
                                        specs.get(f'{key}_version'))
                if log is not None:
                    log.info(f'loading {package} from check point {path}')
                states.append(load_lightning_state_dict(path))
        self.mfiller = MotionInfillerVAE(states[0], self.device)
        self.traj_predictor = TrajPredVAE(states[1], self.device, self.smpl)

    def get_motion_latent(self, seq_len):
        return self.mfiller.get_latent(seq_len)

    def get_traj_latent(self, seq_len):
        return self.traj_predictor.get_latent(seq_len)

    def inference(self, batch, sample_num=1, recon=False):
        """motion_traj_joint_model.py:141-145 (+ pred_trajectory :73-133, 'infer' mode)"""
        if recon:
            raise NotImplementedError('recon mode needs the posterior encoders (training-side, out of scope)')
        data = self.mfiller.inference(batch, sample_num, recon=False, multi_step=True)
        motion = data['infer_out_body_pose']                                        # [B,S,T,69]
        B, S, T = motion.shape[:3]
        tb = {'in_body_pose': motion.reshape(B * S, T, 69)}
        if 'in_traj_latent' in data:
            tb['in_traj_latent'] = data['in_traj_latent']
        out = self.traj_predictor.inference(tb, sample_num=1)
        data['infer_out_pose'] = out['infer_out_pose'].view(B, S, T, 72)
        data['infer_out_trans'] = out['infer_out_trans'].view(B, S, T, 3)
        data['infer_out_orient'] = out['infer_out_orient'].view(B, S, T, 3)
        data['infer_out_local_traj_tp'] = out['infer_out_local_traj_tp'].view(T, B, S, 11)
        return data
