"""``python -m machin_b200.auto {list,generate,launch}`` (surface of machin/auto/__main__.py:12-96).

``launch`` of a sharded framework is run once per rank: ``torchrun --nproc-per-node N -m machin_b200.auto launch --config
c.json`` (the process group is initialised here from the torchrun environment)."""
import argparse
import json
import os
import sys
from pprint import pprint

from .config import (generate_algorithm_config, generate_env_config, generate_training_config, get_available_algorithms,
                     get_available_environments, is_algorithm_distributed, launch)


def _parser():
    parser = argparse.ArgumentParser(prog="python -m machin_b200.auto")
    commands = parser.add_subparsers(dest="command")
    listing = commands.add_parser("list", help="List available algorithms or environments.")
    listing.add_argument("--algo", action="store_true", help="List available algorithms.")
    listing.add_argument("--env", action="store_true", help="List available environments.")
    generate = commands.add_parser("generate", help="Generate configuration.")
    generate.add_argument("--algo", type=str, required=True, help="Algorithm name to use.")
    generate.add_argument("--env", type=str, required=True, help="Environment name to use.")
    generate.add_argument("--print", action="store_true", help="Direct config output to screen.")
    generate.add_argument("--output", type=str, default="config.json", help="JSON config file output path.")
    run = commands.add_parser("launch", help="Launch training.")
    run.add_argument("--config", type=str, default="config.json", help="JSON config file path.")
    return parser


def main(argv=None):
    args = _parser().parse_args(argv)
    if args.command == "list":
        if args.env:
            print("Available environments are:")
            print("\n".join(get_available_environments()))
        elif args.algo:
            print("Available algorithms are:")
            print("\n".join(get_available_algorithms()))
        else:
            print("You can list --algo or --env.")
    elif args.command == "generate":
        if args.algo not in get_available_algorithms():
            print(f"{args.algo} is not a valid algorithm, use list --algo to get a list of available algorithms.")
            return 0
        if args.env not in get_available_environments():
            print(f"{args.env} is not a valid environment, use list --env to get a list of available environments.")
            return 0
        config = generate_training_config(config=generate_algorithm_config(args.algo, config=generate_env_config(args.env, {})))
        if args.print:
            pprint(config)
        with open(args.output, "w") as out:
            json.dump(config, out, indent=4, sort_keys=True)      # (gradient_max = inf is written as Infinity, which json.load reads back)
        print(f"Config saved to {args.output}")
    elif args.command == "launch":
        with open(args.config) as src:
            config = json.load(src)
        if is_algorithm_distributed(config) and "RANK" in os.environ:
            import torch as t
            import torch.distributed as dist
            if not dist.is_initialized():
                on_gpu = t.cuda.is_available()
                if on_gpu:
                    t.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
                dist.init_process_group("nccl" if on_gpu else "gloo")
        launch(config)
    else:
        _parser().print_help()
    return 0


if __name__ == "__main__":
    sys.exit(main())
