"""CLI with the reference's flags and YAML keys (LiDARGen/main.py:17-163,177-209).

    python -m sdpc_b200.main --sample --ni --config Line.yml [--seed 1234 --exp exp --doc doc -i images]

`--config` is resolved against this package's `configs/` (Line.yml, Inpainting.yml, Densification.yml: the
names the reference README uses) or taken as a path.  Like the reference, any exception raised while
sampling is logged with its traceback and the process still returns 0 (main.py:200-209).
"""
import argparse
import logging
import os
import sys
import traceback

import numpy as np
import torch
import yaml


def dict2namespace(config):
    namespace = argparse.Namespace()
    for key, value in config.items():
        setattr(namespace, key, dict2namespace(value) if isinstance(value, dict) else value)
    return namespace


def parse_args_and_config(argv=None):
    parser = argparse.ArgumentParser(description=globals()['__doc__'], formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument('--config', type=str, required=True, help='Path to the config file')
    parser.add_argument('--seed', type=int, default=1234, help='Random seed')
    parser.add_argument('--exp', type=str, default='exp', help='Path for saving running related data.')
    parser.add_argument('--doc', type=str, default='b200', help='Name of the log folder.')
    parser.add_argument('--comment', type=str, default='', help='A string for experiment comment')
    parser.add_argument('--verbose', type=str, default='info', help='Verbose level: info | debug | warning | critical')
    parser.add_argument('--test', action='store_true', help='Whether to test the model (not supported: out of scope)')
    parser.add_argument('--sample', action='store_true', help='Whether to produce samples from the model')
    parser.add_argument('--nvs', action='store_true', help='(not supported: the reference path behind it imports a missing module)')
    parser.add_argument('--fast_fid', action='store_true', help='(not supported: out of scope)')
    parser.add_argument('--resume_training', action='store_true', help='(not supported: out of scope)')
    parser.add_argument('-i', '--image_folder', type=str, default='images', help="The folder name of samples")
    parser.add_argument('--ni', action='store_true', help="No interaction. Suitable for Slurm Job launcher")
    parser.add_argument('--densification', action='store_true', help='densification flag (main.py:48)')
    args = parser.parse_args(argv)
    args.log_path = os.path.join(args.exp, 'logs', args.doc)
    path = args.config
    if not os.path.exists(path):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'configs', args.config)
    with open(path, 'r') as f:
        config = yaml.safe_load(f)
    if "image_width" not in config["data"]:                        # main.py:43-44
        config["data"]["image_width"] = config["data"]["image_size"]
    new_config = dict2namespace(config)
    # forced overrides of the reference (main.py:46-48)
    new_config.sampling.inpainting = True
    new_config.sampling.interpolation = False
    new_config.sampling.densification = args.densification
    if not args.sample:
        raise SystemExit("only --sample is supported: training / test / nvs / fast_fid are outside the hot path (SURVEY.md 8)")
    os.makedirs(os.path.join(args.exp, 'image_samples'), exist_ok=True)
    args.image_folder = os.path.join(args.exp, 'image_samples', args.image_folder)
    if os.path.exists(args.image_folder) and not args.ni:
        if input(f"Image folder {args.image_folder} already exists. Overwrite? (Y/N)").upper() != 'Y':
            raise SystemExit("Output image folder exists. Program halted.")
    os.makedirs(args.image_folder, exist_ok=True)
    level = getattr(logging, args.verbose.upper(), None)
    if not isinstance(level, int):
        raise ValueError('level {} not supported'.format(args.verbose))
    logging.basicConfig(level=level, format='%(levelname)s - %(filename)s - %(asctime)s - %(message)s')
    device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    logging.info("Using device: {}".format(device))
    new_config.device = device
    torch.manual_seed(args.seed)                                    # main.py:156-159
    np.random.seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(args.seed)
    return args, new_config


def main(argv=None):
    args, config = parse_args_and_config(argv)
    logging.info("Writing log file to {}".format(args.log_path))
    logging.info("Exp instance id = {}".format(os.getpid()))
    logging.info("Exp comment = {}".format(args.comment))
    from .runner import NCSNRunnerAllForOne, NCSNRunnerKITTISimultaneous
    try:
        if config.data.dataset == 'KITTI360_im_8batch':             # main.py:191-195
            runner = NCSNRunnerKITTISimultaneous(args, config)
        else:
            runner = NCSNRunnerAllForOne(args, config)
        runner.sample()
    except Exception:
        logging.error(traceback.format_exc())
    return 0


if __name__ == '__main__':
    sys.exit(main())
