"""
Command-line / config front door of the reference's drivers (mirror of src/util/args.py:9-112):
same flags, same expconf.conf lookup (-n <expname> -> conf file, data dir), returns (args, conf).
"""
import argparse
import os

from .conf import ConfigFactory

_PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))


def parse_args(callback=None, training=False, default_conf="conf/default_mv.conf", default_expname="example",
               default_data_format="dvr", default_num_epochs=10000000, default_lr=1e-4, default_gamma=1.00,
               default_datadir="data", default_ray_batch_size=50000, argv=None, make_dirs=True):
    p = argparse.ArgumentParser()
    p.add_argument("--conf", "-c", type=str, default=None)
    p.add_argument("--resume", "-r", action="store_true", help="continue training")
    p.add_argument("--gpu_id", type=str, default="0", help="GPU(s) to use, space delimited")
    p.add_argument("--name", "-n", type=str, default=default_expname, help="experiment name")
    p.add_argument("--dataset_format", "-F", type=str, default=None,
                   help="Dataset format, multi_obj | dvr | dvr_gen | dvr_dtu | srn")
    p.add_argument("--exp_group_name", "-G", type=str, default=None)
    p.add_argument("--logs_path", type=str, default="logs")
    p.add_argument("--checkpoints_path", type=str, default="checkpoints")
    p.add_argument("--visual_path", type=str, default="visuals")
    p.add_argument("--epochs", type=int, default=default_num_epochs)
    p.add_argument("--lr", type=float, default=default_lr)
    p.add_argument("--gamma", type=float, default=default_gamma)
    p.add_argument("--datadir", "-D", type=str, default=None)
    p.add_argument("--ray_batch_size", "-R", type=int, default=default_ray_batch_size)
    if callback is not None:
        p = callback(p)
    args = p.parse_args(argv)
    if args.exp_group_name is not None:
        for k in ("logs_path", "checkpoints_path", "visual_path"):
            setattr(args, k, os.path.join(getattr(args, k), args.exp_group_name))
    if make_dirs:
        os.makedirs(os.path.join(args.checkpoints_path, args.name), exist_ok=True)
        os.makedirs(os.path.join(args.visual_path, args.name), exist_ok=True)
    expconf = ConfigFactory.parse_file(os.path.join(_PROJECT_ROOT, "expconf.conf"))
    if args.conf is None:
        args.conf = expconf.get_string("config." + args.name, default_conf)
    if args.datadir is None:
        args.datadir = expconf.get_string("datadir." + args.name, default_datadir)
    conf_path = args.conf if os.path.isabs(args.conf) or os.path.exists(args.conf) else os.path.join(_PROJECT_ROOT, args.conf)
    conf = ConfigFactory.parse_file(conf_path)
    if args.dataset_format is None:
        args.dataset_format = conf.get_string("data.format", default_data_format)
    args.gpu_id = list(map(int, args.gpu_id.split()))
    print("EXPERIMENT NAME:", args.name)
    if training:
        print("CONTINUE?", "yes" if args.resume else "no")
    print("* Config file:", args.conf)
    print("* Dataset format:", args.dataset_format)
    print("* Dataset location:", args.datadir)
    return args, conf
