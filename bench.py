#!/usr/bin/env python
"""Benchmark of the GLAMR global-optimisation hot path (BASELINE.json metric: global-opt iterations/sec over
frames x persons), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one optimiser iteration of GlobalReconOptimizer.optimize_main (trajectory codec + camera + full SMPL
LBS for every frame-person + projection + residuals + analytic backward [+ one NCCL allreduce of the packed
gradient when N > 1] + Adam).  Workload (config.workload): the glamr_dynamic stage on N persons x 300 frames, the
persons sharded one per GPU (weak scaling; N = 1 is BASELINE.json configs[1]).  Prints ONE JSON line on rank 0.

--impl reference times the reference algorithm's CPU path on this box's host cores: the oracle port under oracle/
(torch-CPU restatement pinned to the executed reference by tests/golden), because /root/reference is not on the
GPU box.  It is the only place besides tests/ and smoke() that executes oracle/ code.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FRAMES = 300
CFG_ID = 'glamr_dynamic'
BYTES_CONST = 19_595_160          # SMPL constants, fp32 dense
BYTES_PER_FP = 1_460              # per frame-person reads + gradient writes + Adam traffic
FLOPS_PER_FP = 15.85e6            # dense full-LBS forward
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
REF_BUDGET_S = float(os.environ.get('GLAMR_REF_BUDGET_S', 150.0))   # wall-clock bound (s) of the CPU reference arm (--impl reference)
# dram__bytes_read.sum + dram__bytes_write.sum of one lbs_kernel launch, keyed by frame-persons per launch (ncu capture, profiles/)
NCU_LBS_DRAM_BYTES = {300: 18942208}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=FRAMES)
    ap.add_argument('--persons', type=int, default=0, help='default: one per GPU')
    ap.add_argument('--lbs-mode', default='full', choices=['full'])
    ap.add_argument('--cpu-sample-iters', type=int, default=12)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_problem(args, persons):
    from glamr_b200.config import Config
    from glamr_b200.synthetic import make_in_dict, make_smpl_assets
    assets = make_smpl_assets(0)
    in_dict = make_in_dict(assets, persons, args.frames, seed=0, seq_name='bench')
    cfg = Config(CFG_ID, out_dir='/tmp/glamr_b200_bench')
    return assets, in_dict, cfg


def cpu_port_timing(assets, in_dict, cfg, iters, warm=3):
    """the oracle port (torch CPU, all host threads) on the same workload: seconds per iteration"""
    import torch
    from glamr_b200.synthetic import LatentInjector
    from glamr_b200.synthetic_nets import make_prior_states
    from oracle.global_opt import OracleGlobalRecon
    from oracle.nets import MotionTrajJoint
    from oracle.smpl import OracleSMPL
    cfg = copy.deepcopy(cfg)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would time a single-threaded reference)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    stage, specs = next(iter(cfg.opt_stage_specs.items()))
    st_m, st_t = make_prior_states(1234)
    model = OracleGlobalRecon(cfg, assets, mt_model=LatentInjector(MotionTrajJoint(st_m, st_t, OracleSMPL(assets)), 0))
    data = model.init_data(copy.deepcopy(in_dict))
    times = []
    model.optimize_main(data, specs['opt_variables'], specs['opt_lr'], warm + iters, specs['loss_cfg'], {'stage': stage},
                        on_iter=lambda it, last, dt: times.append(dt))
    t = np.array(times[warm:])
    return float(np.median(t)), float(t.min()), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import torch
    persons = args.persons or args.gpus
    assets, in_dict, cfg = make_problem(args, persons)
    K, W = max(args.steps, 1), max(args.warmup, 1)
    # Each step is a bounded sample of the workload so that K + W steps end within REF_BUDGET_S: a probe of W iterations on
    # the full workload gives the per-iteration time; if K of them do not fit, a step processes the first `sample` persons
    # only (the reference loops over persons, its cost per frame-person is the same) and the metric counts those units.
    sample = persons
    probe, _, cores = cpu_port_timing(assets, in_dict, cfg, 1, warm=min(W, 2))
    if probe * (K + W) > REF_BUDGET_S and persons > 1:
        sample = max(1, min(persons, int(persons * REF_BUDGET_S / (probe * (K + W)))))
        assets, in_dict, cfg = make_problem(args, sample)
    k_run = K
    if probe * sample / persons * (K + W) > REF_BUDGET_S:          # one person still too slow: fewer timed steps, stated in `sample`
        k_run = max(3, int(REF_BUDGET_S / (probe * sample / persons)) - W)
    med, mn, cores = cpu_port_timing(assets, in_dict, cfg, k_run, warm=W)
    units = sample * args.frames
    val = units / med
    out = {
        'impl': 'reference', 'metric': 'global_opt_frame_person_iterations_per_sec', 'value': val, 'unit': 'frame*person*iter/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'iters_per_sec': 1.0 / med,
        'config': {'workload': f'{CFG_ID}:init_opt, {persons} person(s) x {args.frames} frames, full-LBS every iteration', 'persons': persons,
                   'frames': args.frames},
        'cpu_baseline': {'value': val, 'unit': 'frame*person*iter/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{k_run} timed iterations (median) after {W} warm-up of the oracle port (torch CPU), each over {sample} of the {persons} person(s) x {args.frames} frames'},
        'e2e': {'value': val, 'unit': 'frame*person*iter/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(out))


def nbytes(x):
    import torch
    if isinstance(x, np.ndarray):
        return x.nbytes
    if isinstance(x, torch.Tensor):
        return x.numel() * x.element_size()
    if isinstance(x, dict):
        return sum(nbytes(v) for v in x.values())
    return 0


def _dbg(*a):
    if os.environ.get('BENCH_DEBUG'):
        print(f'[bench rank {os.environ.get("RANK", 0)}] {time.time() % 1000:.1f}', *a, file=sys.stderr, flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from glamr_b200 import lib as L
    from glamr_b200.recon import GlobalReconOptimizer
    from glamr_b200.smpl import SMPL
    from glamr_b200.motion_traj import MotionTrajJointModel
    from glamr_b200.synthetic import LatentInjector
    from glamr_b200.synthetic_nets import make_prior_states
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    persons = args.persons or world
    assets, in_dict, cfg = make_problem(args, persons)
    stage, specs = next(iter(cfg.opt_stage_specs.items()))
    K, W = args.steps, max(args.warmup, 3)
    smpl = SMPL(assets, device=dev)
    _dbg('problem made')
    prior = MotionTrajJointModel(None, dev, None, smpl=smpl, states=make_prior_states(1234))

    def new_model(graph=True):
        c = copy.deepcopy(cfg)
        c.grecon_model_specs['use_cuda_graph'] = graph
        return GlobalReconOptimizer(c, dev, None, smpl=smpl, mt_model=LatentInjector(prior, 0), dist=(rank, world) if world > 1 else None)

    # ---------------- device-resident timing: K iterations, L2 flushed between iterations, CUDA events per iteration
    model = new_model()
    data = model.init_data(copy.deepcopy(in_dict))
    model._cur_vars, model._cur_stage, model._loss_cfg = specs['opt_variables'], stage, specs['loss_cfg']
    model._set_stage(data, specs['opt_variables'], specs['loss_cfg'], stage, reset_adam=True, begin=True)
    hist = torch.zeros((W + 3 * K + 128, L.NUM_TERMS + 1), device=dev)
    lib = model._lib

    def iteration():
        model._backward()
        L.check(lib.glamr_opt_apply(model._opt, L.ptr(model._theta), L.ptr(model._reduce), float(specs['opt_lr']), L.ptr(hist), L.NUM_TERMS + 1,
                                    L.stream_ptr()), 'apply')
    # The timed loops replay ONE captured iteration per step.  With the opt-in peer-memory reduction the capture is the
    # library's own (glamr_opt_iterate); otherwise torch captures backward [+ NCCL all-reduce] + apply, which measured
    # ~12 us per step faster than calling glamr_opt_iterate(n=1) from Python once per step (0.153 vs 0.168 ms flushed).
    native = getattr(model, '_peer_ok', False)
    graph = None
    if native:
        def step():
            L.check(lib.glamr_opt_iterate(model._opt, L.ptr(model._theta), L.ptr(model._reduce), float(specs['opt_lr']), L.ptr(hist),
                                          L.NUM_TERMS + 1, 1, 1, L.stream_ptr()), 'iterate')
        for _ in range(W):
            step()                               # the first call runs eagerly and captures, the rest replay
    else:
        for _ in range(W):
            iteration()
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                iteration()
        except Exception:
            graph = None
            torch.cuda.synchronize()
        step = graph.replay if graph is not None else iteration
    _dbg('warm-up done; native', native, 'torch graph', graph is not None)
    for _ in range(3):
        step()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in evs:
        flush.fill_(1)                       # evict L2 (126 MB) between timed iterations
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    cold_ms = sum(a.elapsed_time(b) for a, b in evs)
    _dbg('cold loop done')
    # back-to-back (L2-warm steady state of the real loop), one event pair around K replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    _dbg('warm loop done')
    t = torch.tensor([cold_ms, warm_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cold_ms, warm_ms = float(t[0]), float(t[1])

    # ---------------- LBS kernel alone (roofline): eager iterations with the library's event pair around the kernel
    L.check(lib.glamr_opt_kernel_timing(model._opt, 1), 'timing')
    import ctypes
    lbs = []
    for _ in range(min(K, 50)):
        flush.fill_(1)
        iteration()
        ms = ctypes.c_float()
        L.check(lib.glamr_opt_last_lbs_ms(model._opt, ctypes.byref(ms)), 'lbs_ms')
        lbs.append(ms.value)
    L.check(lib.glamr_opt_kernel_timing(model._opt, 0), 'timing')
    lbs_ms = float(np.mean(lbs))
    _dbg('lbs timing done')
    n_local = model._n_range[1] - model._n_range[0]

    # ---------------- end to end through the public API with host buffers: optimize(in_dict numpy) -> numpy dict
    e2e_model = new_model()
    c2 = e2e_model.cfg
    for st in c2.opt_stage_specs.values():
        st['opt_niters'] = K
    e2e_model.optimize(copy.deepcopy(in_dict))                      # warm-up call (one-time CUDA/graph setup)
    _dbg('e2e warm-up done')
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = e2e_model.optimize(copy.deepcopy(in_dict))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    _dbg('e2e done')
    tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt[0])
    h2d = nbytes(in_dict['est'])
    d2h = nbytes({k: v for k, v in out.items() if k != 'gt'})

    if rank == 0:
        units = persons * args.frames
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        alg_bytes = BYTES_CONST + n_local * BYTES_PER_FP
        achieved = alg_bytes / (lbs_ms * 1e-3) / 1e9
        fp32_tf = FLOPS_PER_FP * n_local / (lbs_ms * 1e-3) / 1e12
        res = {
            'metric': 'global_opt_frame_person_iterations_per_sec', 'value': units * K / (cold_ms * 1e-3), 'unit': 'frame*person*iter/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': cold_ms / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'iters_per_sec': K / (cold_ms * 1e-3),
            'value_l2_warm': units * K / (warm_ms * 1e-3), 'ms_per_step_l2_warm': warm_ms / K,
            'config': {'workload': f'{CFG_ID}:init_opt, {persons} person(s) x {args.frames} frames, full-LBS every iteration',
                       'persons': persons, 'frames': args.frames, 'parallelism': (f'persons sharded over {world} GPU(s), ' + ('gradient reduction over NVLink peer memory fused into the Adam kernel' if getattr(model, '_peer_ok', False) else '1 NCCL allreduce/iter')) if world > 1 else 'single GPU',
                       'l2': 'flushed between timed iterations (256 MiB fill); value_l2_warm = back-to-back replays',
                       'cuda_graph': bool(native or graph is not None), 'lbs_mode': args.lbs_mode, 'prior': 'CUDA infiller+traj-pred with seeded stand-in weights (no checkpoints offline), latents injected'},
            'clocks': clocks,
            'gpu_launches': 6 * K,
            'e2e': {'value': units * K / e2e_s, 'unit': 'frame*person*iter/s', 'h2d_bytes_per_step': h2d / K, 'd2h_bytes_per_step': d2h / K,
                    'seconds': e2e_s, 'what': f'GlobalReconOptimizer.optimize(in_dict numpy)->numpy dict incl. init_data, {K} iterations',
                    'phase_seconds': {k: round(v, 5) for k, v in e2e_model.phase_seconds.items()},
                    'loop_ms_per_iter': round(e2e_model.iter_ms[-1][2], 5)},
            'roofline': {'bound': 'hbm', 'kernel': 'lbs_kernel', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                         'traffic': NCU_LBS_DRAM_BYTES.get(n_local), 'traffic_source': 'profiles/lbs_kernel_r01_final.md (ncu --set full, dram__bytes_read + write, per launch)' if n_local in NCU_LBS_DRAM_BYTES else None,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                         'algorithmic_bytes': alg_bytes, 'kernel_ms': lbs_ms, 'kernel_share_of_step': lbs_ms / (cold_ms / K),
                         'fp32': {'achieved_tflops': fp32_tf, 'peak_tflops_nominal': FP32_NOMINAL_TFLOPS, 'frac': fp32_tf / FP32_NOMINAL_TFLOPS,
                                  'note': 'the fused full-LBS iteration is FP32-FMA bound with L2-resident constants (SURVEY §8d); HBM fraction is small by construction'}},
        }
        if not args.no_cpu_baseline:
            med, mn, cores = cpu_port_timing(assets, in_dict, cfg, args.cpu_sample_iters)
            res['cpu_baseline'] = {'value': units / med, 'unit': 'frame*person*iter/s', 'cores': cores, 'kind': 'port',
                                   'sample': f'{args.cpu_sample_iters} iterations (median; min {units / mn:.0f}) of the oracle port on the same workload after 3 warm-up'}
        print(json.dumps(res))
    if world > 1:
        # captured graphs hold NCCL work: release them before tearing the process group down, and leave without the
        # interpreter's shutdown path (a destroy with live captures can block)
        del graph, step, model, e2e_model
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
