#!/usr/bin/env python3
"""Ragged all-gather demo for 2 ranks (reference: tutel/examples/nccl_all_gather_v.py).
    torchrun --nproc_per_node=2 -m tutel_b200.examples.nccl_all_gather_v"""
import torch

from tutel_b200 import net, system


def main():
    env = system.init_data_model_parallel(backend='nccl' if torch.cuda.is_available() else 'gloo', group_count=1)
    dev = env.local_device
    assert env.global_size == 2, 'This test case is set for World Size == 2 only'
    data = torch.tensor([10] * 5 if env.global_rank == 0 else [20] * 3, device=dev)
    print('Device-%d sends: %s' % (env.global_rank, [data]))
    net.barrier()
    print('Device-%d recvs: %s' % (env.global_rank, net.batch_all_gather_v([data])[0]))


if __name__ == '__main__':
    main()
