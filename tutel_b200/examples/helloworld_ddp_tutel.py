#!/usr/bin/env python3
"""ZeRO-style ``net.TutelDistributedOptimizer``: shared parameters' optimizer state is sharded over the ranks,
expert parameters step locally (reference: tutel/examples/helloworld_ddp_tutel.py)."""
import torch

from tutel_b200 import net
from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer


def main(argv=None):
    args = base_parser().parse_args(argv)
    s = Session(args)
    layer = default_layer(s, scan_expert_func=None, gate_type={'type': 'top', 'k': args.top, 'fp32_gate': args.fp32_gate})
    s.report_params(layer)
    model = MoEClassifier(layer).to(s.device)
    s.print(model)
    opt = net.TutelDistributedOptimizer(model.parameters(), group=None, average_shared=True).warp_local(torch.optim.SGD, lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner(', parallel_type = `%s`' % args.parallel_type)
    s.train(model, opt, x, y)


if __name__ == '__main__':
    main()
