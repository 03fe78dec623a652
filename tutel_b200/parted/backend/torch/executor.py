"""Runtime half of generated programs: session setup, sharded synthetic tensors, timed execution
(reference: tutel/parted/backend/torch/executor.py:13-115; its missing `simple_all_reduce` import is fixed here)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

from tutel_b200 import net as C
from tutel_b200 import system

parallel_env = None
fusable_params = set()


def warp_bwd_allreduce(data, is_param):
    if is_param:
        fusable_params.add(id(data))
        return C.allreduce_backward(data, group=parallel_env.global_group)
    return C.allreduce_backward(data, group=parallel_env.model_group)


def sharded_randn(shape, dim, dtype, requires_grad=False, is_param=False, device=None):
    device = device if device is not None else parallel_env.local_device
    torch.manual_seed(1)
    full = torch.randn(shape, dtype=dtype, device='cpu').to(device)
    if dim >= 0:
        part = torch.chunk(full, chunks=parallel_env.model_size, dim=dim)[parallel_env.model_rank].contiguous()
    elif dim == -2:
        assert full.numel() % parallel_env.model_size == 0
        part = full.view(parallel_env.model_size, -1)[parallel_env.model_rank].contiguous()
    else:
        part = full.contiguous()
    if is_param:
        part = torch.nn.Parameter(part * 1e-3)
        part.is_param = True
    else:
        part.requires_grad_(requires_grad)
    if dim == -2:
        part._full_shape = shape
        part.is_param = True
    part.dim_state = dim
    return part


def init_session(group_size, group_count=1, device_type='cuda'):
    global parallel_env, fusable_params
    parallel_env = system.init_data_model_parallel(group_count=group_count, backend='nccl' if device_type == 'cuda' else 'gloo')
    fusable_params = set()
    assert parallel_env.model_size == group_size, \
        'This program was generated for %d-way parallelism while the session has %d device(s).' % (group_size, parallel_env.model_size)


def model_executor(module, is_training=True):
    name = module.compute_name
    model = module().to(parallel_env.local_device)
    inputs = module.synthetic_inputs()
    output = model(**inputs)
    params = list(model.parameters())
    verbose = int(os.environ.get('VERBOSE', '0'))
    is_cuda = parallel_env.local_device.type == 'cuda'
    is_training = is_training and isinstance(output, torch.Tensor)
    digest = float(output.contiguous().view(-1)[0]) if isinstance(output, torch.Tensor) else -1
    if is_training:
        torch.manual_seed(1)
        label = torch.LongTensor(output.size(0)).random_(1).to(output.device)
        optimizer = torch.optim.SGD(params, lr=1e-5) if params else None

    def sync():
        if parallel_env.is_distributed and parallel_env.global_size > 1:
            dist.barrier()
        if is_cuda:
            torch.cuda.synchronize(parallel_env.local_device)

    def one_step():
        sync()
        t0 = time.time()
        if is_training:
            if optimizer:
                optimizer.zero_grad()
            out = model(**inputs).contiguous()
            loss = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(out.view(out.size(0), -1), dim=1), label)
            loss.backward()
            if parallel_env.group_count > 1:
                for p in params:
                    if id(p) not in fusable_params and p.grad is not None:
                        p.grad = C.simple_all_reduce(p.grad, group=parallel_env.data_group)
            if optimizer:
                optimizer.step()
        else:
            with torch.no_grad():
                model(**inputs)
        sync()
        return time.time() - t0

    for _ in range(5):
        one_step()
    step_time = sum(one_step() for _ in range(5)) / 5
    if parallel_env.global_rank == 0:
        if verbose:
            sys.stderr.write('  [%s] digest = %g .., time = %g\n' % (name, digest, step_time))
        result = json.dumps({'name': name, 'step_time': step_time, 'digest': digest})
        if 'CONFIG_STORE_PATH' in os.environ:
            with open(os.environ['CONFIG_STORE_PATH'], 'w') as f:
                f.write(result)
        print(result)
