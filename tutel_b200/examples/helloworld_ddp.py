#!/usr/bin/env python3
"""torch DistributedDataParallel for the shared parameters, expert parameters excluded through
``_ddp_params_and_buffers_to_ignore`` (reference: tutel/examples/helloworld_ddp.py)."""
import torch

from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer


def main(argv=None):
    args = base_parser().parse_args(argv)
    s = Session(args)
    layer = default_layer(s)
    s.report_params(layer)
    model = MoEClassifier(layer).to(s.device)
    if s.world > 1:
        model._ddp_params_and_buffers_to_ignore = [n for n, p in model.named_parameters() if hasattr(p, 'skip_allreduce')]
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[s.device.index] if s.device.type == 'cuda' else None)
        ddp._moe_layer = model._moe_layer
        model = ddp
    s.print(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner(', parallel_type = `%s`' % args.parallel_type)
    s.train(model, opt, x, y)


if __name__ == '__main__':
    main()
