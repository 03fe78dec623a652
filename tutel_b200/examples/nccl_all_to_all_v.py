#!/usr/bin/env python3
"""Ragged all-to-all demo for 2 ranks (reference: tutel/examples/nccl_all_to_all_v.py).
    torchrun --nproc_per_node=2 -m tutel_b200.examples.nccl_all_to_all_v"""
import torch

from tutel_b200 import net, system


def main():
    env = system.init_data_model_parallel(backend='nccl' if torch.cuda.is_available() else 'gloo', group_count=1)
    dev = env.local_device
    assert env.global_size == 2, 'This test case is set for World Size == 2 only'
    if env.global_rank == 0:
        data, counts = torch.tensor([10, 10, 10, 10, 10], device=dev), torch.tensor([1, 4], dtype=torch.int64, device=dev)
    else:
        data, counts = torch.tensor([20, 20, 20], device=dev), torch.tensor([2, 1], dtype=torch.int64, device=dev)
    print('Device-%d sends: %s' % (env.global_rank, [data]))
    net.barrier()
    print('Device-%d recvs: %s' % (env.global_rank, net.batch_all_to_all_v([data], counts)[0]))


if __name__ == '__main__':
    main()
