"""Evaluation sweep: the reference's ``global_recon/eval_dataset.py`` on the CUDA Evaluator -- for every sequence and seed
load ``<results_dir>/<seq>/grecon/<seq>_seed<n>.pkl`` (the optimiser's output pickle, with ``gt`` / ``gt_meta`` inside),
compute the per-sequence metrics, reduce over seeds (eval_dataset.py:41-54) and print the accumulated line.

    python -m glamr_b200.global_recon.eval_dataset --dataset 3dpw --results_dir out/3dpw --seeds 1,2,3
    python -m glamr_b200.global_recon.eval_dataset --sequences synthetic_0000,synthetic_0001 --results_dir out/sweep

``--sequences`` overrides the reference's hard-coded 3DPW test list (all 24 names; the reference's loop stops after the first
two, ``sequences[:2]`` at eval_dataset.py:41 -- ``--limit 2`` reproduces that)."""
import argparse
import pickle

import torch

TEST_SEQUENCES = {
    '3dpw': ['downtown_arguing_00', 'downtown_bar_00', 'downtown_bus_00', 'downtown_cafe_00', 'downtown_car_00', 'downtown_crossStreets_00',
             'downtown_downstairs_00', 'downtown_enterShop_00', 'downtown_rampAndStairs_00', 'downtown_runForBus_00', 'downtown_runForBus_01',
             'downtown_sitOnStairs_00', 'downtown_stairs_00', 'downtown_upstairs_00', 'downtown_walkBridge_01', 'downtown_walkUphill_00',
             'downtown_walking_00', 'downtown_warmWelcome_00', 'downtown_weeklyMarket_00', 'downtown_windowShopping_00', 'flat_guitar_01',
             'flat_packBags_00', 'office_phoneCall_00', 'outdoors_fencing_01'],
}


def run(args, make_evaluator=None):
    from glamr_b200.evaluator import Evaluator
    seeds = [int(x) for x in str(args.seeds).split(',')]
    multi = len(seeds) > 1
    sequences = args.sequences.split(',') if args.sequences else TEST_SEQUENCES[args.dataset]
    if args.limit:
        sequences = sequences[:args.limit]
    device = torch.device('cuda', args.gpu)
    torch.cuda.set_device(device)
    torch.set_grad_enabled(False)
    mk = make_evaluator or (lambda log_file: Evaluator(args.results_dir, args.dataset, device=device, log_file=log_file, compute_sample=multi))
    evaluator, seed_evaluator = mk(f'{args.results_dir}/log_eval.txt'), mk(f'{args.results_dir}/log_eval_seed.txt')
    for sind, seq_name in enumerate(sequences):
        arr = []
        evaluator.log.info(f'{sind}/{len(sequences)} evaluating global reconstruction for {seq_name}')
        for seed in seeds:
            with open(f'{args.results_dir}/{seq_name}/grecon/{seq_name}_seed{seed}.pkl', 'rb') as f:
                data = pickle.load(f)
            arr.append(seed_evaluator.compute_sequence_metrics(data, seq_name, accumulate=False))
        all_seeds = evaluator.metrics_from_multiple_seeds(arr)
        evaluator.update_accumulated_metrics(all_seeds, seq_name)
        evaluator.print_metrics(all_seeds, prefix=f'{sind}/{len(sequences)} --- All seeds {seq_name} --- ', print_accum=False)
    evaluator.print_metrics(prefix='Total ------- ', print_accum=True)
    return evaluator


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--dataset', default='3dpw')
    ap.add_argument('--results_dir', default='out/3dpw')
    ap.add_argument('--gpu', type=int, default=0)
    ap.add_argument('--seeds', default='1')
    ap.add_argument('--sequences', default='')
    ap.add_argument('--limit', type=int, default=0)
    return ap.parse_args(argv)


if __name__ == '__main__':
    run(parse())
