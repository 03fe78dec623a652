#!/usr/bin/env python3
"""A custom expert whose single weight is flat-sharded over the GPUs that share an expert and re-assembled with
``net.zero_gather`` every forward (reference: tutel/examples/helloworld_custom_expert_sharded.py).

    torchrun --nproc_per_node=2 -m tutel_b200.examples.helloworld_custom_expert_sharded --num_local_experts=-2
"""
import torch

from tutel_b200 import net
from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer, manual_allreduce


class ShardedLinearExpert(torch.nn.Module):
    def __init__(self, model_dim, num_experts_per_device, sharded_count, my_config=None):
        super().__init__()
        self.sharded_count = sharded_count
        self.full_shape = torch.Size([num_experts_per_device, model_dim, model_dim])
        shard = (self.full_shape.numel() + sharded_count - 1) // sharded_count
        self.W = torch.nn.Parameter(torch.empty(shard).normal_(0, 0.001))
        self.act = torch.nn.functional.relu if my_config == 'relu' else None

    def forward(self, x, ctx):
        group = net.create_groups_from_world(group_count=-self.sharded_count, parent_group=ctx.group).model_group
        w = net.zero_gather(self.W, full_shape=self.full_shape, group=group)
        y = torch.matmul(x, w)
        return self.act(y) if self.act is not None else y


def main(argv=None):
    args = base_parser().parse_args(argv)
    s = Session(args)
    layer = default_layer(s, parallel_type='adaptive:1',
                          experts={'type': 'custom', 'module': ShardedLinearExpert,
                                   'num_experts_per_device': args.num_local_experts, 'my_config': None})
    s.report_params(layer)
    model = MoEClassifier(layer).to(s.device)
    s.print(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner()
    s.train(model, opt, x, y, sync_grads=manual_allreduce(s, model))


if __name__ == '__main__':
    main()
