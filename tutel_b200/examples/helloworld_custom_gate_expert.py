#!/usr/bin/env python3
"""Plugging a custom gate and a custom expert module into the layer (reference:
tutel/examples/helloworld_custom_gate_expert.py)."""
import torch

from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer, manual_allreduce


class CustomGate(torch.nn.Module):
    def __init__(self, model_dim, num_global_experts, k=1, **options):
        super().__init__()
        self.top_k = min(num_global_experts, int(k))
        self.wg = torch.nn.Parameter(torch.randn(model_dim, num_global_experts) * 1e-3)

    def forward(self, x):
        return torch.matmul(x, self.wg.to(x.dtype))


class CustomExpert(torch.nn.Module):
    def __init__(self, model_dim, num_experts_per_device, sharded_count, hidden_size=2048, my_config=None):
        super().__init__()
        assert sharded_count == 1, 'this demo expert keeps whole experts on one device'
        self.w1 = torch.nn.Parameter(torch.randn(num_experts_per_device, model_dim, hidden_size) * 1e-3)
        self.w2 = torch.nn.Parameter(torch.randn(num_experts_per_device, hidden_size, model_dim) * 1e-3)
        self.act = torch.nn.functional.gelu if my_config == 'gelu' else torch.nn.functional.relu

    def forward(self, x, ctx):
        return torch.matmul(self.act(torch.matmul(x, self.w1)), self.w2)


def main(argv=None):
    args = base_parser().parse_args(argv)
    s = Session(args)
    layer = default_layer(
        s, gate_type={'type': 'custom', 'module': CustomGate, 'k': args.top},
        experts={'type': 'custom', 'module': CustomExpert, 'num_experts_per_device': args.num_local_experts,
                 'hidden_size': args.hidden_size, 'my_config': None})
    s.report_params(layer)
    model = MoEClassifier(layer).to(s.device)
    s.print(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner()
    s.train(model, opt, x, y, sync_grads=manual_allreduce(s, model))


if __name__ == '__main__':
    main()
