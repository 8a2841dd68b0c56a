#!/usr/bin/env python
"""CLI with the reference's flags (/root/reference src/distilation/main.py:9-27): -lt / -ct / -k / -ch / -r."""
import argparse

from . import config, lstm_train, mlp_train


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-lt", "--lstm_train", help="train lstm", action="store_true")
    parser.add_argument("-ct", "--mlp_train", help="train mlp", action="store_true")
    parser.add_argument("-k", "--keep_prob", help=" keep_prob on ob dropout ", nargs=1, default=None)
    parser.add_argument("-ch", "--check", help="check point", action="store_true")
    parser.add_argument("-r", "--restore", help="restore", action="store_true")
    parser.add_argument("--num_envs", type=int, default=config.NUM_ENVS)
    parser.add_argument("--iterations", type=int, default=None)
    args = parser.parse_args(argv)
    keep_prob = float(args.keep_prob[0]) if args.keep_prob else config.KEEP_PROB
    if args.check:
        import os
        import torch
        path = os.path.join(config.base_path, "student_mlp_b200.pt")
        print(" checking saved variables ")
        sd = torch.load(path)
        def shapes(d):
            return {k: (shapes(v) if isinstance(v, dict) else tuple(v.shape) if hasattr(v, "shape") else v) for k, v in d.items()}
        print(shapes(sd))
    elif args.lstm_train:
        lstm_train.train(True, args.restore, num_envs=args.num_envs, iterations=args.iterations, keep_prob=keep_prob)
    elif args.mlp_train:
        mlp_train.train(True, args.restore, num_envs=args.num_envs, iterations=args.iterations, keep_prob=keep_prob)


if __name__ == "__main__":
    main()
