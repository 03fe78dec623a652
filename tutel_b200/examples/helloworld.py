#!/usr/bin/env python3
"""MoE layer benchmark / regression driver.  Same command line and log format as tutel/examples/helloworld.py, so the
reference's golden loss curves (tests/test_baseline.json) remain a drop-in oracle; built on the shared example session
(:mod:`tutel_b200.examples._driver`) like every other hello-world program.

    python -m tutel_b200.examples.helloworld --batch_size=16
    python -m torch.distributed.run --nproc_per_node=8 -m tutel_b200.examples.helloworld --dtype bfloat16 ...
"""
import os

import torch
import torch.nn.functional as F

from tutel_b200 import system
from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer, manual_allreduce


def build_parser():
    p = base_parser()
    p.add_argument('--megablocks_size', type=int, default=0)        # > 0: dropless block-sparse expert path (inference)
    p.add_argument('--use_tensorcore', default=False, action='store_true')
    p.add_argument('--expert_type', type=str, default='ffn')
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.use_tensorcore:
        torch.backends.cuda.matmul.allow_tf32 = True
    s = Session(args)

    layer = default_layer(s, experts={'type': args.expert_type, 'num_experts_per_device': args.num_local_experts,
                                      'hidden_size_per_expert': args.hidden_size, 'activation_fn': lambda t: F.relu(t)})
    s.report_params(layer)
    call = (lambda moe, x: moe(x, megablocks_size=args.megablocks_size)) if args.megablocks_size > 0 else None
    model = MoEClassifier(layer, call).to(s.device)
    s.print(model)

    ckpt = None
    if args.checkpoint_path:
        ckpt = system.apply_rank_size_from_pattern(args.checkpoint_path, rank=s.rank, size=s.world)
        if os.path.exists(ckpt):
            model.load_state_dict(torch.load(ckpt))
        else:
            print('Checkpoint not loaded: file `%s` is not found. Will train the model from start.' % ckpt)

    optimizer = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner(extra=', parallel_type = `%s`' % args.parallel_type)
    s.train(model, optimizer, x, y, sync_grads=manual_allreduce(s, model))

    if ckpt:
        torch.save(model.state_dict(), ckpt)


if __name__ == '__main__':
    main()
