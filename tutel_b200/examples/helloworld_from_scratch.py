#!/usr/bin/env python3
"""Building an MoE block from the low-level ops: ``moe.top_k_routing`` + ``moe.fast_encode`` + ``net.all_to_all`` +
``moe.fast_decode`` (reference: tutel/examples/helloworld_from_scratch.py)."""
import torch
import torch.nn.functional as F

from tutel_b200 import moe, net, system


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--device', type=str, default='cuda' if torch.cuda.is_available() else 'cpu')
    ap.add_argument('--num_steps', type=int, default=10)
    ap.add_argument('--model_dim', type=int, default=2048)
    ap.add_argument('--hidden_size', type=int, default=2048)
    ap.add_argument('--num_samples', type=int, default=4096)
    args = ap.parse_args()
    env = system.init_data_model_parallel(backend='nccl' if args.device == 'cuda' else 'gloo')
    M, H, El = args.model_dim, args.hidden_size, 2
    E = El * env.global_size

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(1)
            self.wg = torch.nn.Parameter(torch.randn(M, E) * 1e-3)
            torch.manual_seed(env.global_rank + 1)
            self.w1 = torch.nn.Parameter(torch.randn(El, M, H) * 1e-3)
            self.w2 = torch.nn.Parameter(torch.randn(El, H, M) * 1e-3)
            self.b1 = torch.nn.Parameter(torch.zeros(El, 1, H))
            self.b2 = torch.nn.Parameter(torch.zeros(El, 1, M))
            for p in (self.w1, self.w2, self.b1, self.b2):
                p.skip_allreduce = True

        def forward(self, x, k=2):
            scores = F.softmax(torch.matmul(x, self.wg), dim=-1)
            crit, l_aux = moe.top_k_routing(scores, top_k=k)
            y = moe.fast_encode(x, crit)              # [E, C, M]
            y = net.all_to_all(y, 1, 0)               # [El, W*C, M]
            y = torch.matmul(F.relu(torch.matmul(y, self.w1) + self.b1), self.w2) + self.b2
            y = net.all_to_all(y, 0, 1)               # [E, C, M]
            return moe.fast_decode(y, crit), l_aux

    model = Block().to(env.local_device)
    torch.manual_seed(env.global_rank + 1)
    data = torch.randn([args.num_samples, M], device=env.local_device)
    label = torch.LongTensor(args.num_samples).random_(1).to(env.local_device)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    for i in range(args.num_steps):
        t0 = system.record_time()
        opt.zero_grad()
        out, l_aux = model(data)
        loss = F.nll_loss(F.log_softmax(out, dim=1), label) + 0.0001 * l_aux
        loss.backward()
        for p in model.parameters():
            if not hasattr(p, 'skip_allreduce'):
                p.grad = net.simple_all_reduce(p.grad)
        opt.step()
        env.dist_print('STEP-%d: loss = %.5f, step_time = %.3f s' % (i, loss, system.record_time() - t0))


if __name__ == '__main__':
    main()
