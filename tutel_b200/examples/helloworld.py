#!/usr/bin/env python3
"""MoE layer benchmark / regression driver: same CLI and output format as tutel/examples/helloworld.py, so the
reference's golden loss curves (tests/test_baseline.json) remain a drop-in oracle.

    python -m tutel_b200.examples.helloworld --batch_size=16
    python -m torch.distributed.run --nproc_per_node=8 -m tutel_b200.examples.helloworld --dtype bfloat16 ...
"""
import argparse
import os

import torch
import torch.nn.functional as F

from tutel_b200 import moe as tutel_moe
from tutel_b200 import net, system


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--local_rank', type=int, default=-1)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--num_tokens', type=int, default=512)
    p.add_argument('--model_dim', type=int, default=2048)
    p.add_argument('--hidden_size', type=int, default=2048)
    p.add_argument('--num_local_experts', type=int, default=2)
    p.add_argument('--dtype', type=str, default='float32')
    p.add_argument('--fp32_gate', default=False, action='store_true')
    p.add_argument('--top', type=int, default=2)
    p.add_argument('--l_aux_wt', type=float, default=0.0)
    p.add_argument('--a2a_ffn_overlap_degree', type=int, default=1)
    p.add_argument('--allreduce_degree', type=int, default=1)
    p.add_argument('--num_steps', type=int, default=100)
    p.add_argument('--parallel_type', type=str, default='adaptive:1')
    p.add_argument('--checkpoint_path', type=str, default='')
    p.add_argument('--device', type=str, default='cuda' if torch.cuda.is_available() else 'cpu')
    p.add_argument('--use_2dh', default=False, action='store_true')
    p.add_argument('--eval', default=False, action='store_true')
    p.add_argument('--capacity_factor', type=float, default=1.0)  # 0.0: dropless, negative: no-padded capacity
    p.add_argument('--megablocks_size', type=int, default=0)
    p.add_argument('--use_tensorcore', default=False, action='store_true')
    p.add_argument('--expert_type', type=str, default='ffn')
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.use_tensorcore:
        torch.backends.cuda.matmul.allow_tf32 = True

    env = system.init_data_model_parallel(backend='nccl' if args.device == 'cuda' else 'gloo')
    rank, world, dist_print = env.global_rank, env.global_size, env.dist_print
    device = env.local_device
    args.local_rank = device.index

    dtypes = {'float32': torch.float32, 'float64': torch.float64, 'float16': torch.float16, 'bfloat16': torch.bfloat16}
    if args.dtype not in dtypes:
        raise Exception('Unrecognized data type specified: %s' % args.dtype)
    torch.set_default_dtype(dtypes[args.dtype])

    class ExampleModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._moe_layer = tutel_moe.moe_layer(
                gate_type={'type': 'top', 'k': args.top, 'fp32_gate': args.fp32_gate, 'capacity_factor': args.capacity_factor},
                experts={'type': args.expert_type, 'num_experts_per_device': args.num_local_experts,
                         'hidden_size_per_expert': args.hidden_size, 'activation_fn': lambda x: F.relu(x)},
                model_dim=args.model_dim,
                scan_expert_func=lambda name, param: setattr(param, 'skip_allreduce', True),
                seeds=(1, rank + 1, 1),
                a2a_ffn_overlap_degree=args.a2a_ffn_overlap_degree,
                parallel_type=args.parallel_type,
                use_2dh=args.use_2dh,
            )
            local_count = sum(p.numel() for _, p in self._moe_layer.get_parameter_iterator(param_type='local_experts'))
            shared_count = sum(p.numel() for _, p in self._moe_layer.get_parameter_iterator(param_type='gate'))
            dist_print('[Statistics] param count for MoE local_experts = %s, param count for MoE gate = %s.\n' % (local_count, shared_count))

        def forward(self, input):
            if args.megablocks_size > 0:
                result = self._moe_layer(input, megablocks_size=args.megablocks_size)
            else:
                result = self._moe_layer(input)
            return F.log_softmax(torch.sum(result, dim=2), dim=1)

    model = ExampleModel().to(device)
    dist_print(model)

    checkpoint_path = None
    if args.checkpoint_path:
        checkpoint_path = system.apply_rank_size_from_pattern(args.checkpoint_path, rank=rank, size=world)
        if os.path.exists(checkpoint_path):
            model.load_state_dict(torch.load(checkpoint_path))
        else:
            print('Checkpoint not loaded: file `%s` is not found. Will train the model from start.' % checkpoint_path)

    optimizer = torch.optim.SGD(model.parameters(), lr=1e-5)

    torch.manual_seed(0)
    x = torch.randn([args.batch_size, args.num_tokens, args.model_dim], dtype=torch.float32, device='cpu')
    x = x.to(dtype=torch.get_default_dtype(), device=device)
    y = torch.LongTensor(args.batch_size).random_(1).to(device)

    dist_print('[Benchmark] world_size = %s, dtype = %s, model_dim = %s, hidden_size = %s, samples = %s, num_local_experts = %s, topK = %s, a2a_ffn_overlap_degree = %s, parallel_type = `%s`, device = `%s`' % (
        world, args.dtype, args.model_dim, args.hidden_size, args.batch_size * args.num_tokens, args.num_local_experts,
        args.top, args.a2a_ffn_overlap_degree, args.parallel_type, device))

    if args.allreduce_degree == -1:
        params_for_all_reduce = []
    else:
        params_for_all_reduce = [p for p in model.parameters() if not hasattr(p, 'skip_allreduce') and getattr(p, 'requires_grad', False)]

    average_time, num_steps = 0, args.num_steps
    num_global_experts = tutel_moe.moe_layer.global_expert_count(args.num_local_experts, group=system.get_local_session().model_group)
    for i in range(num_steps):
        t_start = system.record_time()
        if not args.eval:
            optimizer.zero_grad()
            output = model(x)
            loss = F.nll_loss(output, y)
            if args.l_aux_wt:
                loss += args.l_aux_wt * model._moe_layer.l_aux
            loss.backward()
            if world > 1:
                for p in params_for_all_reduce:
                    p.grad /= world
                    p.grad = net.simple_all_reduce(p.grad)
            optimizer.step()
        else:
            with torch.no_grad():
                output = model(x)
                loss = F.nll_loss(output, y)
        t_stop = system.record_time()

        mm_ceof, cap_ceof = 1 if args.eval else 3, min(args.top, num_global_experts)
        tflops = (args.batch_size * args.num_tokens * args.model_dim * args.hidden_size) * 4 * mm_ceof * cap_ceof * 1e-12 / (t_stop - t_start)
        dist_print('STEP-%s: loss = %.5f, step_time = %.6f sec, perf = %.2f tflops.' % (i, float(loss.data), t_stop - t_start, tflops))
        if i + 10 >= num_steps:
            average_time += t_stop - t_start

    average_time /= 10
    dist_print('\n[Summary] Average synchronized step_time = %s sec.' % average_time)

    if checkpoint_path:
        torch.save(model.state_dict(), checkpoint_path)


if __name__ == '__main__':
    main()
