#!/usr/bin/env python
"""Command-line driver with the reference's interface (hidasib/GRU4Rec run.py:10-133): same flags, same
parameter-file / parameter-string formats, same printed lines (paropt.py parses `PRIMARY METRIC:`).
The model class is loaded through the reference's plugin seam `-g GRFILE` (default: the root-level `gru4rec`
module, i.e. the B200 implementation)."""
import argparse
import importlib
import importlib.util
import os
import sys
import time
from collections import OrderedDict


def build_parser():
    p = argparse.ArgumentParser(description='Train or load a GRU4Rec model and measure recall / MRR on test set(s).')
    p.add_argument('path', metavar='PATH', type=str, help='Training data (TAB separated .tsv/.txt or pickled DataFrame .pickle), or the serialized model when --load_model is given.')
    p.add_argument('-ps', '--parameter_string', metavar='PARAM_STRING', type=str, help='Training parameters as `name1=value1,name2=value2`; booleans True/False; lists use / (e.g. layers=200/200). Exclusive with -pf and -l.')
    p.add_argument('-pf', '--parameter_file', metavar='PARAM_PATH', type=str, help='Python file defining an OrderedDict named `gru4rec_params`. Exclusive with -ps and -l.')
    p.add_argument('-l', '--load_model', action='store_true', help='Load a trained model from PATH instead of training. Exclusive with -ps and -pf.')
    p.add_argument('-s', '--save_model', metavar='MODEL_PATH', type=str, help='Save the trained model to MODEL_PATH.')
    p.add_argument('-t', '--test', metavar='TEST_PATH', type=str, nargs='+', help='Test data set(s).')
    p.add_argument('-m', '--measure', metavar='AT', type=int, nargs='+', default=[20], help='Recommendation list length(s) for recall & MRR (default: 20).')
    p.add_argument('-e', '--eval_type', metavar='EVAL_TYPE', choices=['standard', 'conservative', 'median', 'tiebreaking'], default='standard', help='Tie handling of the ranking (see evaluate_gpu).')
    p.add_argument('-ss', '--sample_store_size', metavar='SS', type=int, default=10000000, help='Size of the negative-sample buffer in ids (default: 10000000).')
    p.add_argument('--sample_store_on_cpu', action='store_true', help='Legacy: draw the negative samples on the host.')
    p.add_argument('-g', '--gru4rec_model', metavar='GRFILE', type=str, default='gru4rec', help='Module that provides the GRU4Rec class (default: gru4rec).')
    p.add_argument('-ik', '--item_key', metavar='IK', type=str, default='ItemId', help='Item id column (default: ItemId).')
    p.add_argument('-sk', '--session_key', metavar='SK', type=str, default='SessionId', help='Session id column (default: SessionId).')
    p.add_argument('-tk', '--time_key', metavar='TK', type=str, default='Time', help='Timestamp column (default: Time).')
    p.add_argument('-pm', '--primary_metric', metavar='METRIC', choices=['recall', 'mrr'], default='recall', help='Primary metric for -lpm (default: recall).')
    p.add_argument('-lpm', '--log_primary_metric', action='store_true', help='Print `PRIMARY METRIC: value` at the end (one test file, one list length).')
    return p


def load_data(fname, args):
    """TSV (ItemId read as str, SessionId as int32) or pickled DataFrame, with the reference's column checks (run.py:45-78)."""
    import pandas as pd
    import joblib
    keys = [('session IDs', args.session_key, 'SessionId', 'session_key'), ('item IDs', args.item_key, 'ItemId', 'item_key'), ('time', args.time_key, 'Time', 'time_key')]
    if fname.endswith('.pickle'):
        print('Loading data from pickle file: {}'.format(fname))
        data = joblib.load(fname)
        columns = list(data.columns)
    else:
        with open(fname, 'rt') as f:
            columns = f.readline().strip().split('\t')
        data = None
    for what, key, default, pname in keys:
        if key not in columns:
            print('ERROR. The column specified for {} "{}" is not in the data file ({})'.format(what, key, fname))
            print('The default column name is "{}", but you can specify otherwise by setting the `{}` parameter of the model.'.format(default, pname))
            sys.exit(1)
    if data is None:
        print('Loading data from TAB separated file: {}'.format(fname))
        data = pd.read_csv(fname, sep='\t', usecols=[args.session_key, args.item_key, args.time_key], dtype={args.session_key: 'int32', args.item_key: 'str'})
    return data


def main(argv=None):
    args = build_parser().parse_args(argv)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    GRU4Rec = importlib.import_module(args.gru4rec_model).GRU4Rec
    import evaluation
    if (args.parameter_string is not None) + (args.parameter_file is not None) + (args.load_model) != 1:
        print('ERROR. Exactly one of the following parameters must be provided: --parameter_string, --parameter_file, --load_model')
        sys.exit(1)
    if args.load_model:
        print('Loading trained model from file: {}'.format(args.path))
        gru = GRU4Rec.loadmodel(args.path)
    else:
        if args.parameter_file:
            param_file_path = os.path.abspath(args.parameter_file)
            spec = importlib.util.spec_from_file_location(os.path.split(param_file_path)[1].split('.py')[0], param_file_path)
            params = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(params)
            gru4rec_params = params.gru4rec_params
            print('Loaded parameters from file: {}'.format(param_file_path))
        if args.parameter_string:
            gru4rec_params = OrderedDict([x.split('=') for x in args.parameter_string.split(',')])
        print('Creating GRU4Rec model')
        gru = GRU4Rec()
        gru.set_params(**gru4rec_params)
        print('Loading training data...')
        data = load_data(args.path, args)
        store_type = 'cpu' if args.sample_store_on_cpu else 'gpu'
        if store_type == 'cpu':
            print('WARNING! The sample store is set to be on the CPU. This will make training significantly slower on the GPU.')
        print('Started training')
        t0 = time.time()
        gru.fit(data, sample_store=args.sample_store_size, store_type=store_type)
        t1 = time.time()
        print('Total training time: {:.2f}s'.format(t1 - t0))
        if args.save_model is not None:
            print('Saving trained model to: {}'.format(args.save_model))
            gru.savemodel(args.save_model)
    if args.test is not None:
        pm_index = {'recall': 0, 'mrr': 1}[args.primary_metric.lower()]
        for test_file in args.test:
            print('Loading test data...')
            test_data = load_data(test_file, args)
            print('Starting evaluation (cut-off={}, using {} mode for tiebreaking)'.format(args.measure, args.eval_type))
            t0 = time.time()
            res = evaluation.evaluate_gpu(gru, test_data, batch_size=512, cut_off=args.measure, mode=args.eval_type, item_key=args.item_key, session_key=args.session_key, time_key=args.time_key)
            t1 = time.time()
            print('Evaluation took {:.2f}s'.format(t1 - t0))
            for i, c in enumerate(args.measure):
                print('Recall@{}: {:.6f} MRR@{}: {:.6f}'.format(c, res[0][i], c, res[1][i]))
            if args.log_primary_metric:
                print('PRIMARY METRIC: {}'.format(res[pm_index][0]))


if __name__ == '__main__':
    main()
