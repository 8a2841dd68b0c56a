#!/usr/bin/env python
"""CLI with the reference's flags (/root/reference src/distilation/main.py:9-27): -lt / -ct / -k / -ch / -r."""
import argparse

from . import config, lstm2_train, lstm_train, mlp_train


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-lt", "--lstm_train", help="train lstm", action="store_true")
    parser.add_argument("-ct", "--mlp_train", help="train mlp", action="store_true")
    parser.add_argument("-k", "--keep_prob", help=" keep_prob on ob dropout ", nargs=1, default=None)
    parser.add_argument("-ch", "--check", help="check point", action="store_true")
    parser.add_argument("-r", "--restore", help="restore", action="store_true")
    parser.add_argument("--num_envs", type=int, default=config.NUM_ENVS)
    parser.add_argument("--two_headed", action="store_true", help="with -lt: the two-headed LSTM graph (action + reward heads, loss = KL + squared reward "
                        "error) of backup/student_rollout.py:130-200,328 instead of student_nn.student_lstm_graph")
    parser.add_argument("--iterations", type=int, default=None)
    parser.add_argument("--teacher", default=None, help=".npz with the teacher weights teacher.py:17-20 restores from teacher.ckpt (flat `params` or the "
                        "baselines variables pi/obfilter/*, pi/pol/*); default: <base_path>/teacher.npz, else a seeded UNTRAINED teacher (announced)")
    parser.add_argument("--checkpoint", default=None, help="loop checkpoint written at the end of -ct / every episode of -lt, read by -r and -ch "
                        "(default: <base_path>/student_mlp_b200.pt, student_lstm_b200.pt)")
    args = parser.parse_args(argv)
    keep_prob = float(args.keep_prob[0]) if args.keep_prob else config.KEEP_PROB
    if args.check:
        import os
        import torch
        default = "student_lstm_b200.pt" if args.lstm_train else "student_mlp_b200.pt"
        path = args.checkpoint or os.path.join(config.base_path, default)
        print(" checking saved variables ")
        sd = torch.load(path)
        def shapes(d):
            return {k: (shapes(v) if isinstance(v, dict) else tuple(v.shape) if hasattr(v, "shape") else v) for k, v in d.items()}
        out = shapes(sd)
        print(out)
        return out
    elif args.lstm_train and args.two_headed:
        return lstm2_train.lstm_train(True, keep_prob, args.checkpoint, args.restore, num_envs=args.num_envs, iterations=args.iterations,
                                      teacher_ckpt=args.teacher)
    elif args.lstm_train:
        return lstm_train.train(True, args.restore, num_envs=args.num_envs, iterations=args.iterations, keep_prob=keep_prob, teacher_ckpt=args.teacher,
                                checkpoint=args.checkpoint)
    elif args.mlp_train:
        return mlp_train.train(True, args.restore, num_envs=args.num_envs, iterations=args.iterations, keep_prob=keep_prob, teacher_ckpt=args.teacher,
                               checkpoint=args.checkpoint)


if __name__ == "__main__":
    main()
