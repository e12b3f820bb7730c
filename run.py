#!/usr/bin/env python
"""Command-line driver with the interface of the reference's run.py (hidasib/GRU4Rec run.py:10-133): same flags, same
parameter-file / parameter-string formats, same printed lines (paropt.py parses `PRIMARY METRIC:`).  The model class comes
in through the reference's plugin seam `-g GRFILE` (default: the root-level `gru4rec` module = the B200 implementation)."""
import argparse
import importlib
import importlib.util
import os
import sys
import time
from collections import OrderedDict

# (flags, keyword arguments) of every command-line option; names, defaults and choices follow run.py:11-26
_TIE_MODES = ['standard', 'conservative', 'median', 'tiebreaking']
_OPTIONS = [
    (('path',), dict(metavar='PATH', type=str,
                     help='Training data (TAB separated .tsv/.txt or pickled DataFrame .pickle), or the serialized model when --load_model is given.')),
    (('-ps', '--parameter_string'), dict(metavar='PARAM_STRING', type=str,
                                         help='Training parameters as `name1=value1,name2=value2`; booleans True/False; lists use / (e.g. layers=200/200). Exclusive with -pf and -l.')),
    (('-pf', '--parameter_file'), dict(metavar='PARAM_PATH', type=str,
                                       help='Python file defining an OrderedDict named `gru4rec_params`. Exclusive with -ps and -l.')),
    (('-l', '--load_model'), dict(action='store_true', help='Load a trained model from PATH instead of training. Exclusive with -ps and -pf.')),
    (('-s', '--save_model'), dict(metavar='MODEL_PATH', type=str, help='Save the trained model to MODEL_PATH.')),
    (('-t', '--test'), dict(metavar='TEST_PATH', type=str, nargs='+', help='Test data set(s).')),
    (('-m', '--measure'), dict(metavar='AT', type=int, nargs='+', default=[20], help='Recommendation list length(s) for recall & MRR (default: 20).')),
    (('-e', '--eval_type'), dict(metavar='EVAL_TYPE', choices=_TIE_MODES, default='standard', help='Tie handling of the ranking (see evaluate_gpu).')),
    (('-ss', '--sample_store_size'), dict(metavar='SS', type=int, default=10000000, help='Size of the negative-sample buffer in ids (default: 10000000).')),
    (('--sample_store_on_cpu',), dict(action='store_true', help='Legacy: draw the negative samples on the host.')),
    (('-g', '--gru4rec_model'), dict(metavar='GRFILE', type=str, default='gru4rec', help='Module that provides the GRU4Rec class (default: gru4rec).')),
    (('-ik', '--item_key'), dict(metavar='IK', type=str, default='ItemId', help='Item id column (default: ItemId).')),
    (('-sk', '--session_key'), dict(metavar='SK', type=str, default='SessionId', help='Session id column (default: SessionId).')),
    (('-tk', '--time_key'), dict(metavar='TK', type=str, default='Time', help='Timestamp column (default: Time).')),
    (('-pm', '--primary_metric'), dict(metavar='METRIC', choices=['recall', 'mrr'], default='recall', help='Primary metric for -lpm (default: recall).')),
    (('-lpm', '--log_primary_metric'), dict(action='store_true', help='Print `PRIMARY METRIC: value` at the end (one test file, one list length).')),
]


def build_parser():
    parser = argparse.ArgumentParser(description='Train or load a GRU4Rec model and measure recall / MRR on test set(s).')
    for flags, kwargs in _OPTIONS:
        parser.add_argument(*flags, **kwargs)
    return parser


def _abort(*lines):
    for line in lines:
        print(line)
    sys.exit(1)


def load_data(fname, args):
    """TSV (item ids read as str, session ids as int32) or pickled DataFrame; the three key columns must exist (run.py:45-78)."""
    import pandas as pd
    import joblib
    pickled = fname.endswith('.pickle')
    if pickled:
        print('Loading data from pickle file: {}'.format(fname))
        frame = joblib.load(fname)
        present = list(frame.columns)
    else:
        with open(fname, 'rt') as handle:
            present = handle.readline().strip().split('\t')
    required = (('session IDs', args.session_key, 'SessionId', 'session_key'),
                ('item IDs', args.item_key, 'ItemId', 'item_key'),
                ('time', args.time_key, 'Time', 'time_key'))
    for role, column, default_name, param_name in required:
        if column not in present:
            _abort('ERROR. The column specified for {} "{}" is not in the data file ({})'.format(role, column, fname),
                   'The default column name is "{}", but you can specify otherwise by setting the `{}` parameter of the model.'.format(default_name, param_name))
    if not pickled:
        print('Loading data from TAB separated file: {}'.format(fname))
        frame = pd.read_csv(fname, sep='\t', usecols=[args.session_key, args.item_key, args.time_key],
                            dtype={args.session_key: 'int32', args.item_key: 'str'})
    return frame


def _training_parameters(args):
    """OrderedDict of constructor parameters from -pf (a Python file defining `gru4rec_params`) or -ps (name=value,...)."""
    if args.parameter_file:
        location = os.path.abspath(args.parameter_file)
        module_name = os.path.split(location)[1].split('.py')[0]
        spec = importlib.util.spec_from_file_location(module_name, location)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        print('Loaded parameters from file: {}'.format(location))
        return module.gru4rec_params
    return OrderedDict(pair.split('=') for pair in args.parameter_string.split(','))


def _train(model_class, args):
    params = _training_parameters(args)
    print('Creating GRU4Rec model')
    gru = model_class()
    gru.set_params(**params)
    print('Loading training data...')
    frame = load_data(args.path, args)
    store_type = 'cpu' if args.sample_store_on_cpu else 'gpu'
    if args.sample_store_on_cpu:
        print('WARNING! The sample store is set to be on the CPU. This will make training significantly slower on the GPU.')
    print('Started training')
    started = time.time()
    gru.fit(frame, sample_store=args.sample_store_size, store_type=store_type)
    print('Total training time: {:.2f}s'.format(time.time() - started))
    if args.save_model is not None and getattr(args, 'rank', 0) == 0:
        print('Saving trained model to: {}'.format(args.save_model))
        gru.savemodel(args.save_model)
    return gru


def _evaluate(gru, evaluation, args):
    primary = ('recall', 'mrr').index(args.primary_metric.lower())
    for test_file in args.test:
        print('Loading test data...')
        frame = load_data(test_file, args)
        print('Starting evaluation (cut-off={}, using {} mode for tiebreaking)'.format(args.measure, args.eval_type))
        started = time.time()
        result = evaluation.evaluate_gpu(gru, frame, batch_size=512, cut_off=args.measure, mode=args.eval_type,
                                         item_key=args.item_key, session_key=args.session_key, time_key=args.time_key)
        print('Evaluation took {:.2f}s'.format(time.time() - started))
        for position, cut in enumerate(args.measure):
            print('Recall@{}: {:.6f} MRR@{}: {:.6f}'.format(cut, result[0][position], cut, result[1][position]))
        if args.log_primary_metric:
            print('PRIMARY METRIC: {}'.format(result[primary][0]))


def _join_distributed_job():
    """`torchrun --nproc-per-node N run.py ...` (one process per GPU; the reference is single-device): join the job the launcher
    described, so that fit() trains data-parallel and evaluate_gpu() scores a shard of the test sessions per rank.  Every rank
    runs the same command on the same files; only rank 0 prints and saves.  Returns (world_size, rank)."""
    if int(os.environ.get('WORLD_SIZE', '1') or 1) <= 1:
        return 1, 0
    from gru4rec_b200.parallel import init_from_env
    world, rank = init_from_env()
    if rank != 0:
        sys.stdout = open(os.devnull, 'w')
    return world, rank


def main(argv=None):
    args = build_parser().parse_args(argv)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    world, rank = _join_distributed_job()
    args.rank = rank
    model_class = importlib.import_module(args.gru4rec_model).GRU4Rec
    import evaluation
    chosen = [args.parameter_string is not None, args.parameter_file is not None, bool(args.load_model)]
    if sum(chosen) != 1:
        _abort('ERROR. Exactly one of the following parameters must be provided: --parameter_string, --parameter_file, --load_model')
    if args.load_model:
        print('Loading trained model from file: {}'.format(args.path))
        gru = model_class.loadmodel(args.path)
    else:
        gru = _train(model_class, args)
    if args.test is not None:
        _evaluate(gru, evaluation, args)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
