#!/usr/bin/env python3
"""Collective bandwidth test (reference: tutel/examples/bandwidth_test.py), with device-side timing.

Prints, per collective, both the reference's figure (total bytes / time) and the bus bandwidth
``bytes_per_rank * (W-1)/W / t`` measured with CUDA events (max over ranks), for the NVLink P2P kernels and - with
``--compare_nccl`` - for NCCL.  ``--sweep`` runs 1 KB ... 1 GB (BASELINE config #5).

    torchrun --nproc_per_node=8 -m tutel_b200.examples.bandwidth_test --size_mb=256
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from tutel_b200 import net, system


def timed(fn, loops, device, is_cuda):
    for _ in range(3):
        fn()
    if is_cuda:
        torch.cuda.synchronize()
        dist.barrier() if dist.is_initialized() and dist.get_world_size() > 1 else None
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(loops):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) * 1e-3 / loops], device=device, dtype=torch.float64)
    else:
        t0 = time.perf_counter()
        for _ in range(loops):
            fn()
        t = torch.tensor([(time.perf_counter() - t0) / loops], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--device', type=str, default='cuda' if torch.cuda.is_available() else 'cpu')
    ap.add_argument('--size_mb', type=int, default=256)
    ap.add_argument('--loop', type=int, default=20)
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--compare_nccl', action='store_true')
    ap.add_argument('--json', type=str, default='')
    args = ap.parse_args(argv)
    env = system.init_data_model_parallel(backend='nccl' if args.device == 'cuda' else 'gloo')
    W, dev, is_cuda = env.global_size, env.local_device, env.local_device.type == 'cuda'
    sizes = [1 << p for p in range(10, 31, 2)] if args.sweep else [args.size_mb << 20]
    results = []
    for nbytes in sizes:
        n = max(W, nbytes // 4 // W * W)
        x = torch.randn([n], device=dev, dtype=torch.float32)
        row = {'bytes': n * 4, 'world': W}
        with torch.no_grad():
            cases = [('all_to_all', lambda: net.simple_all_to_all(x.view(W, -1))),
                     ('all_gather', lambda: net.simple_all_gather(x.view(W, -1)[env.global_rank])),
                     ('all_reduce', lambda: net.simple_all_reduce(x.view(-1), inplace=True)),
                     ('reduce_scatter', lambda: net.simple_reduce_scatter(x.view(W, -1)))]
            if args.compare_nccl and is_cuda and W > 1:
                out = torch.empty_like(x)
                cases.append(('nccl_all_to_all', lambda: dist.all_to_all_single(out, x)))
                cases.append(('nccl_all_reduce', lambda: dist.all_reduce(out)))
            for name, fn in cases:
                t = timed(fn, args.loop, dev, is_cuda)
                algo = n * 4 * 1e-9 / t
                bus = algo * (W - 1) / W if W > 1 else algo
                row[name] = {'seconds': t, 'GBps_reference_formula': algo, 'bus_GBps': bus}
                env.dist_print('%-16s %12d B across %d rank(s): %.4f GB/s (bus %.4f GB/s, %.1f us)' % (name, n * 4, W, algo, bus, t * 1e6))
        results.append(row)
        env.dist_print('')
    if args.json and env.global_rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, 'w') as f:
            json.dump(results, f, indent=1)


if __name__ == '__main__':
    main()
