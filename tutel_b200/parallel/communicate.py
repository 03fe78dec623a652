"""Process groups and collective primitives (with and without autograd).

API parity with the reference's ``tutel/impls/communicate.py`` (cited per function), re-organised around one
generic *split-and-concat* all-to-all and two interchangeable transports:

* ``torch.distributed`` (NCCL on CUDA, Gloo on CPU) - bootstrap, fallback and the baseline to beat;
* the in-kernel NVLink peer-to-peer transport of :mod:`tutel_b200.parallel.p2p` (symmetric heap + flag protocol),
  used automatically for CUDA tensors when the group spans the whole single-node world.
"""
from __future__ import annotations

import datetime
import logging
import os
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TUTEL_GLOBAL_TIMEOUT_SEC = int(os.environ.get('TUTEL_GLOBAL_TIMEOUT_SEC', 86400))
TUTEL_SKIP_A2A = int(os.environ.get('SKIP_A2A', 0)) > 0

_GROUP_CACHE: dict = {}
_BACKEND: List[Optional[str]] = [None]


# ----------------------------------------------------------------------------------------------------------------
# world queries (reference: communicate.py:20-35)
# ----------------------------------------------------------------------------------------------------------------
def _dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size(group=None) -> int:
    if not _dist_ready():
        return 1
    try:
        return dist.get_world_size(group)
    except Exception:
        return 1


def get_world_rank(group=None) -> int:
    if not _dist_ready():
        return 0
    try:
        return dist.get_rank(group)
    except Exception:
        return 0


def barrier(group=None) -> None:
    if get_world_size(group) > 1:
        dist.barrier(group=group)


def create_standalone_group():
    """A group that contains only the calling rank (reference: communicate.py:43-47)."""
    try:
        return dist.new_group(ranks=[get_world_rank()])
    except Exception:
        return None


class DistributedProperties:
    """Result of :func:`create_groups_from_world` (field names as in communicate.py:120-146)."""

    def __repr__(self):
        keys = ('global_size', 'global_rank', 'group_count', 'data_rank', 'model_size', 'model_rank', 'local_device')
        return 'DistributedProperties(%s)' % ', '.join('%s=%s' % (k, getattr(self, k, None)) for k in keys)


def _maybe_init_process_group(backend: str) -> Tuple[bool, int]:
    """Join the job described by torchrun / OpenMPI environment variables.  Returns (is_distributed, local_rank)."""
    timeout = datetime.timedelta(seconds=TUTEL_GLOBAL_TIMEOUT_SEC)
    env = os.environ
    ompi = 'LOCAL_RANK' not in env and 'OMPI_COMM_WORLD_SIZE' in env
    if _dist_ready():
        local = int(env.get('OMPI_COMM_WORLD_LOCAL_RANK' if ompi else 'LOCAL_RANK', 0))
        return True, local
    if ompi:
        if backend:
            dist.init_process_group(backend=backend, timeout=timeout,
                                    init_method='tcp://%s:%s' % (env['MASTER_ADDR'], env.get('MASTER_PORT', '23456')),
                                    rank=int(env['OMPI_COMM_WORLD_RANK']), world_size=int(env['OMPI_COMM_WORLD_SIZE']))
        return _dist_ready(), int(env['OMPI_COMM_WORLD_LOCAL_RANK'])
    if 'RANK' in env and 'WORLD_SIZE' in env and backend:
        if backend == 'nccl' and torch.cuda.is_available():
            # bind the device before NCCL initialises so that eager connection setup targets the right GPU
            torch.cuda.set_device(min(int(env.get('LOCAL_RANK', 0)), torch.cuda.device_count() - 1))
        dist.init_process_group(backend=backend, timeout=timeout)
        return True, int(env.get('LOCAL_RANK', 0))
    return False, 0


def create_groups_from_world(group_count: int, include_init: Optional[str] = None, parent_group=None):
    """Split the world into ``group_count`` data-parallel replicas of a model-parallel group.

    ``model_group`` = consecutive ranks, ``data_group`` = strided ranks; negative ``group_count`` means "model groups
    of that size".  Reference: communicate.py:49-168.
    """
    parent_size, world_size = get_world_size(parent_group), get_world_size()
    if 1 < parent_size < world_size:
        assert include_init is None, 'Torch distributed environment had been initialized.'
        raise Exception('Splitting nesting groups from a subgroup is yet not allowed.')

    if include_init:
        if _BACKEND[0] is not None:
            assert _BACKEND[0] == include_init, 'Only 1 backend type is allowed, get: %s v.s. %s' % (_BACKEND[0], include_init)
        _BACKEND[0] = include_init
    backend = _BACKEND[0]

    if group_count in _GROUP_CACHE:
        return _GROUP_CACHE[group_count]

    is_distributed, local_rank = _maybe_init_process_group(include_init or '')
    if is_distributed:
        glob_size, glob_rank = dist.get_world_size(), dist.get_rank()
    else:
        glob_size, glob_rank, local_rank = 1, 0, 0

    requested = group_count
    if group_count < 0:
        group_count = glob_size // -group_count
    assert group_count > 0 and glob_size % group_count == 0, \
        'Expected to evenly divide devices into %s groups, while the world size of current session is %s.' % (group_count, glob_size)

    model_size = glob_size // group_count
    model_rank, data_rank = glob_rank % model_size, glob_rank // model_size

    model_group = data_group = global_group = None
    if is_distributed:
        model_group = data_group = global_group = dist.group.WORLD
        timeout = datetime.timedelta(seconds=TUTEL_GLOBAL_TIMEOUT_SEC)
        if model_size != glob_size:
            for gr in range(group_count):
                g = dist.new_group(ranks=list(range(gr * model_size, (gr + 1) * model_size)), timeout=timeout)
                if gr == data_rank:
                    model_group = g
        if group_count != glob_size:
            for mr in range(model_size):
                g = dist.new_group(ranks=list(range(mr, glob_size, model_size)), timeout=timeout)
                if mr == model_rank:
                    data_group = g

    res = DistributedProperties()
    res.global_size, res.global_rank = glob_size, glob_rank
    res.group_count, res.data_rank = group_count, data_rank
    res.model_size, res.model_rank = model_size, model_rank
    if backend == 'nccl':
        ndev = max(torch.cuda.device_count(), 1)
        res.local_device = torch.device('cuda', min(local_rank, ndev - 1))
        torch.cuda.set_device(res.local_device)
    elif backend == 'gloo':
        res.local_device = torch.device('cpu')
    elif backend is None:
        res.local_device = None
    else:
        raise Exception('Unsupported backend type: %s' % backend)
    res.data_group, res.model_group, res.global_group = data_group, model_group, global_group
    res.is_distributed = is_distributed
    res.dist_print = (lambda *a, **k: print(*a, **k) if glob_rank == 0 else None)
    _GROUP_CACHE[requested] = res
    return res


# ----------------------------------------------------------------------------------------------------------------
# transport selection
# ----------------------------------------------------------------------------------------------------------------
def _p2p(group, tensor: torch.Tensor):
    """The NVLink P2P transport for this group, or None when NCCL/Gloo must be used."""
    if not tensor.is_cuda:
        return None
    from . import p2p
    return p2p.transport_for(group)


# ----------------------------------------------------------------------------------------------------------------
# collectives without autograd (reference: communicate.py:173-223)
# ----------------------------------------------------------------------------------------------------------------
def simple_all_reduce(input: torch.Tensor, group=None, op=dist.ReduceOp.SUM, inplace: bool = False):
    if get_world_size(group) == 1:
        return input
    t = _p2p(group, input)
    if t is not None and t.supports_reduce(input, op):
        # one kernel, out of place unless asked otherwise: no defensive clone of the input is needed
        return t.all_reduce_(input, op) if inplace else t.all_reduce(input.contiguous(), op)
    output = input if inplace else input.clone(memory_format=torch.contiguous_format)
    dist.all_reduce(output, op=op, group=group)
    return output


def simple_all_to_all(input: torch.Tensor, group=None, background: bool = False, _alias_ok: bool = False):
    world_size = get_world_size(group)
    input = input.contiguous()
    if world_size == 1 or TUTEL_SKIP_A2A:
        return input if not background else (input, lambda *a: None)
    t = _p2p(group, input)
    if t is not None and not background and input.numel() % world_size == 0:
        # `_alias_ok`: the caller copies the result out of the (re-used) staging area itself
        return t.all_to_all(input, copy=not _alias_ok)
    output = torch.empty_like(input)
    if background:
        work = dist.all_to_all_single(output, input, group=group, async_op=True)
        return output, work.wait
    dist.all_to_all_single(output, input, group=group)
    return output


def simple_split(input: torch.Tensor, group=None):
    world_size = get_world_size(group)
    if world_size == 1:
        return input
    assert input.size(0) % world_size == 0, 'Cannot evenly divide dim length %s into %s slices' % (input.size(0), world_size)
    return input.contiguous().chunk(world_size, dim=0)[get_world_rank(group)]


def simple_reduce_scatter(input: torch.Tensor, group=None, op=dist.ReduceOp.SUM):
    world_size = get_world_size(group)
    if world_size == 1:
        return input
    input = input.contiguous()
    assert input.size(0) % world_size == 0, 'Cannot evenly divide dim length %s into %s slices' % (input.size(0), world_size)
    if not input.is_cuda:
        return simple_split(simple_all_reduce(input, group, op=op), group=group)
    t = _p2p(group, input)
    if t is not None and t.supports_reduce_scatter(input, op):
        return t.reduce_scatter(input, op)
    output = torch.empty_like(input[: input.size(0) // world_size])
    dist.reduce_scatter_tensor(output, input, op=op, group=group)
    return output


def simple_all_gather(input: torch.Tensor, group=None):
    world_size = get_world_size(group)
    if world_size == 1:
        return input
    input = input.contiguous()
    t = _p2p(group, input)
    if t is not None:
        output = t.all_gather(input)
    elif input.is_cuda:
        output = torch.empty([world_size * input.numel()], device=input.device, dtype=input.dtype)
        dist.all_gather_into_tensor(output, input.view(-1), group=group)
    else:
        output = torch.empty([world_size, input.numel()], device=input.device, dtype=input.dtype)
        dist.all_gather(list(output.unbind(0)), input.view(-1), group=group)
    return output.view([-1] + list(input.shape[1:]))


# ----------------------------------------------------------------------------------------------------------------
# ragged collectives (reference: communicate.py:225-255, custom_kernel.cpp:463-518)
# ----------------------------------------------------------------------------------------------------------------
def batch_all_to_all_v(datas: Sequence[torch.Tensor], partition_sizes, group=None):
    """Every tensor of ``datas`` is cut by the same per-peer element counts; returns (outputs, received counts)."""
    assert isinstance(datas, (tuple, list)), 'data type for batch_all_to_all_v() is not a list of tensors'
    world_size = get_world_size(group)
    device = datas[0].device
    in_sizes = partition_sizes.to(torch.int64) if isinstance(partition_sizes, torch.Tensor) else \
        torch.tensor(partition_sizes, dtype=torch.int64, device=device)
    assert in_sizes.numel() == world_size
    if world_size == 1:
        return list(datas), in_sizes
    in_sizes = in_sizes.to(device)
    # every rank learns the whole W x W count matrix in one small all-gather: receive offsets can then be computed
    # by the SENDER, which is what lets the payload be pushed with a single kernel (no second size exchange)
    matrix_t = simple_all_gather(in_sizes.view(1, -1), group=group).view(world_size, world_size)
    matrix = [[int(v) for v in row] for row in matrix_t.tolist()]
    rank = get_world_rank(group)
    out_sizes = matrix_t[:, rank].contiguous()
    datas = [d.contiguous().view(-1) for d in datas]
    t = _p2p(group, datas[0])
    in_list, out_list = matrix[rank], [matrix[s][rank] for s in range(world_size)]
    outputs = []
    for d in datas:
        if t is not None:
            outputs.append(t.all_to_all_v(d, matrix))
        else:
            out = torch.empty([sum(out_list)], dtype=d.dtype, device=d.device)
            dist.all_to_all_single(out, d[: sum(in_list)], output_split_sizes=out_list, input_split_sizes=in_list, group=group)
            outputs.append(out)
    return outputs, out_sizes


def batch_all_gather_v(datas: Sequence[torch.Tensor], group=None):
    """All-gather of tensors whose length differs per rank; returns (outputs, per-rank lengths)."""
    assert isinstance(datas, (tuple, list)), 'data type for batch_all_gather_v() is not a list of tensors'
    datas = [d.contiguous().view(-1) for d in datas]
    device = datas[0].device
    input_size = torch.tensor([int(datas[0].numel())], dtype=torch.int64, device=device)
    world_size = get_world_size(group)
    if world_size == 1:
        return list(datas), input_size
    output_sizes = simple_all_gather(input_size, group=group)
    sizes = [int(v) for v in output_sizes.tolist()]
    t = _p2p(group, datas[0])
    outputs = []
    for d in datas:
        if t is not None:
            outputs.append(t.all_gather_v(d, sizes))
        else:
            # equal-size gather of padded pieces (Gloo and older NCCL builds reject ragged lists), then trim
            width = max(sizes)
            padded = torch.zeros([width], dtype=d.dtype, device=d.device)
            padded[: d.numel()] = d
            pieces = [torch.empty([width], dtype=d.dtype, device=d.device) for _ in sizes]
            dist.all_gather(pieces, padded, group=group)
            outputs.append(torch.cat([p[:n] for p, n in zip(pieces, sizes)]))
    return outputs, output_sizes


# ----------------------------------------------------------------------------------------------------------------
# the generic all-to-all:  split `output_dim` into W pieces (piece p -> rank p), concatenate the W received pieces
# (source-rank major) along `input_dim`.   Reference: PrimAllToAll.transform, communicate.py:432-503.
# ----------------------------------------------------------------------------------------------------------------
def _a2a_pack(x: torch.Tensor, output_dim: int, world: int) -> torch.Tensor:
    shape = list(x.shape)
    assert shape[output_dim] % world == 0, 'dim %d (%d) is not divisible by the group size %d' % (output_dim, shape[output_dim], world)
    v = x.reshape(shape[:output_dim] + [world, shape[output_dim] // world] + shape[output_dim + 1:])
    return v.movedim(output_dim, 0).contiguous()


def _a2a_unpack(y: torch.Tensor, input_dim: int) -> torch.Tensor:
    # y: [W(src), *rest]  ->  src axis moved in front of `input_dim`, then merged with it
    v = y.movedim(0, input_dim)
    shape = list(v.shape)
    return v.reshape(shape[:input_dim] + [shape[input_dim] * shape[input_dim + 1]] + shape[input_dim + 2:])


def _hier_sizes(world: int) -> Tuple[int, int]:
    local = int(os.environ.get('LOCAL_SIZE', 0)) or (torch.cuda.device_count() if torch.cuda.is_available() else 1)
    local = max(1, min(local, world))
    if world % local != 0:
        local = 1
    return world // local, local


_WARNED_2DH = [False]


def _raw_all_to_all(packed: torch.Tensor, group, use_2dh: bool) -> torch.Tensor:
    """All-to-all of a [W, ...] tensor along dim 0."""
    world = get_world_size(group)
    if not use_2dh:
        return simple_all_to_all(packed, group, _alias_ok=True)
    nnodes, ngpus = _hier_sizes(world)
    if nnodes == 1 or ngpus == 1 or (group is not None and group is not dist.group.WORLD and world != get_world_size()):
        # One NVSwitch domain (or one GPU per node): the 2-D hierarchical algorithm degenerates to the flat
        # exchange, exactly like the reference's native fallback (custom_kernel.cpp:681,722-737).
        if not _WARNED_2DH[0] and nnodes == 1 and get_world_rank() == 0:
            logging.info('use_2dh: single NVLink domain detected, using the flat peer-to-peer all-to-all')
            _WARNED_2DH[0] = True
        return simple_all_to_all(packed, group)
    # 2DH: (1) intra-node exchange groups traffic per destination-local-rank, (2) inter-node exchange delivers it; the
    # re-ordering between the phases is a block transpose of whole records (native stride-copy kernel on CUDA).
    env = create_groups_from_world(-ngpus)
    intra, inter = env.model_group, env.data_group
    rest = list(packed.shape[1:])
    x = _block_transpose(packed.reshape([nnodes, ngpus] + rest), nnodes, ngpus)      # [dst_local, dst_node, ...]
    x = simple_all_to_all(x, intra)                                                   # [src_local, dst_node, ...]
    x = _block_transpose(x.reshape([ngpus, nnodes] + rest), ngpus, nnodes)            # [dst_node, src_local, ...]
    x = simple_all_to_all(x, inter)                                                   # [src_node, src_local, ...]
    return x.reshape([world] + rest)


def _block_transpose(x: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """``[rows, cols, ...] -> [cols, rows, ...]`` as a contiguous tensor (one launch of the stride-copy kernel on CUDA)."""
    if x.is_cuda and x.numel() > 0:
        from ..ops import backend
        if backend.has_cuda_ext():
            src = x.contiguous()
            dst = torch.empty([cols, rows] + list(x.shape[2:]), dtype=x.dtype, device=x.device)
            backend.count_launch()
            backend.require_ext().p2p_stride_copy(src, dst, rows, cols)
            return dst
    return x.transpose(0, 1).contiguous()


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, input_dim: int, output_dim: int, group, use_2dh: bool):
        ctx.dims, ctx.group, ctx.use_2dh = (input_dim, output_dim), group, use_2dh
        world = get_world_size(group)
        raw = _raw_all_to_all(_a2a_pack(x, output_dim, world), group, use_2dh)
        out = _a2a_unpack(raw, input_dim)
        if out.is_cuda and out.untyped_storage().data_ptr() == raw.untyped_storage().data_ptr():
            t = _p2p(group, out)
            if t is not None and t.owns(raw):
                out = out.clone()          # still a view of the P2P staging area: detach it before the next collective
        return out

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):
        input_dim, output_dim = ctx.dims
        return _AllToAll.apply(dy.contiguous(), output_dim, input_dim, ctx.group, ctx.use_2dh), None, None, None, None


def all_to_all(input: torch.Tensor, input_dim: int, output_dim: int, group=None, background: bool = False,
               use_2dh: bool = False):
    """Flexible all-to-all with autograd: ``[HY] X LY Z -> [HX] HY LX LY Z``.

    With ``background=True`` the exchange is started asynchronously and a zero-argument callable returning the
    result is handed back (reference: communicate.py:447-486); gradients still flow through the result.
    """
    world = get_world_size(group)
    if background:
        assert not use_2dh, 'Background mode for AllToAll 2DH is not implemented.'
        if input_dim == output_dim or world == 1:
            return lambda *a: input
        packed = _a2a_pack(input, output_dim, world)
        raw, wait = simple_all_to_all(packed, group, background=True)

        def finish(*_):
            wait()
            return _a2a_unpack(_RestoreBackward.apply(raw, packed, group), input_dim)
        return finish
    if input_dim == output_dim or world == 1:
        return input
    return _AllToAll.apply(input, input_dim, output_dim, group, use_2dh)


class _RestoreBackward(torch.autograd.Function):
    """Re-attach an asynchronously produced buffer to the autograd graph of its source (communicate.py:400-410)."""

    @staticmethod
    def forward(ctx: Any, output: torch.Tensor, source: torch.Tensor, group):
        ctx.group = group
        return output.view_as(output)

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):
        return None, simple_all_to_all(grad.contiguous(), ctx.group), None


class _AllToAllSingle(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, group):
        ctx.group = group
        return simple_all_to_all(x, group)

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):
        return _AllToAllSingle.apply(dy, ctx.group), None


def all_to_all_single(input: torch.Tensor, group=None):
    return _AllToAllSingle.apply(input, group)


# ----------------------------------------------------------------------------------------------------------------
# all-reduce / gather / scatter with autograd (reference: communicate.py:505-604)
# ----------------------------------------------------------------------------------------------------------------
class _BwdAllreduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x, op, group):
        ctx.op, ctx.group = op, group
        return x.view_as(x)

    @staticmethod
    def backward(ctx: Any, dy):
        return simple_all_reduce(dy, group=ctx.group, op=ctx.op), None, None


class _FwdAllreduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x, op, group):
        return simple_all_reduce(x, group=group, op=op)

    @staticmethod
    def backward(ctx: Any, dy):
        return dy, None, None


def allreduce_backward(input, op=dist.ReduceOp.SUM, group=None):
    """Identity in forward, all-reduce of the gradient in backward."""
    return _BwdAllreduce.apply(input, op, group)


def allreduce_forward(input, op=dist.ReduceOp.SUM, group=None):
    """All-reduce in forward, identity in backward."""
    return _FwdAllreduce.apply(input, op, group)


def _swap(t: torch.Tensor, a: int, b: int) -> torch.Tensor:
    return t if a == b else t.swapaxes(a, b)


class _ReduceScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x, group):
        ctx.group = group
        return simple_reduce_scatter(x, group)

    @staticmethod
    def backward(ctx: Any, dy):
        return simple_all_gather(dy, ctx.group), None


def reduce_scatter(input, dim: int, group=None):
    return _swap(_ReduceScatter.apply(_swap(input, 0, dim), group), 0, dim)


class _AllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x, fused: bool, group):
        ctx.group, ctx.fused = group, fused
        return simple_all_gather(x, group)

    @staticmethod
    def backward(ctx: Any, dy):
        if ctx.fused:
            return simple_reduce_scatter(dy, ctx.group), None, None
        return simple_split(dy, ctx.group), None, None


def all_gather(input, dim: int, fused: bool = False, group=None):
    """Concatenate along ``dim``; backward slices (``fused=False``) or reduce-scatters (``fused=True``)."""
    return _swap(_AllGather.apply(_swap(input, 0, dim), fused, group), 0, dim)


def zero_gather(input, full_shape=None, group=None):
    """ZeRO-style parameter gather: concat shards on dim 0 (grad: reduce-scatter), drop padding, reshape."""
    if not full_shape:
        full_shape = list(input.shape)
        full_shape[0] *= get_world_size(group)
    numel = 1
    for v in full_shape:
        numel *= int(v)
    out = _AllGather.apply(input, True, group) if get_world_size(group) > 1 else input
    if out.numel() == numel:        # no padding: a pure view (a slice would cost a zero-fill + copy in backward)
        return out.view(full_shape)
    return out.reshape(-1)[:numel].view(full_shape)


def zero_scatter(input, scatter_fn: Callable, group=None):
    """Pad the flattened tensor to a multiple of the group size and apply ``scatter_fn`` (split / reduce-scatter)."""
    group_size = get_world_size(group)
    full = input.numel()
    if full % group_size == 0:
        data = input.reshape(-1)
    else:
        data = torch.zeros([(full + group_size - 1) // group_size * group_size], device=input.device, dtype=input.dtype)
        data[:full] = input.reshape(-1)
    return scatter_fn(data, group=group), input.shape


class _SpatialSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, x, group):
        ctx.group = group
        return simple_split(x, group)

    @staticmethod
    def backward(ctx: Any, dy):
        return simple_all_gather(dy, ctx.group), None


def spatial_split(input, dim: int, group=None):
    return _swap(_SpatialSplit.apply(_swap(input, 0, dim), group), 0, dim)


# ----------------------------------------------------------------------------------------------------------------
# layout helpers around the expert computation (reference: communicate.py:606-622)
# ----------------------------------------------------------------------------------------------------------------
def pre_expert_permute(input: torch.Tensor, group=None):
    """[W*El, C, M] (source-rank major) -> [El, W*C, M]."""
    world = get_world_size(group)
    if world == 1:
        return input
    v = input.view([world, -1] + list(input.shape[1:]))
    v = v.transpose(0, 1).contiguous()
    return v.view([v.shape[0], -1] + list(v.shape[3:]))


def post_expert_permute(input: torch.Tensor, group=None):
    """[El, W*C, M'] -> [W*El, C, M']."""
    world = get_world_size(group)
    if world == 1:
        return input
    v = input.view([input.shape[0], world, -1] + list(input.shape[2:]))
    v = v.transpose(0, 1).contiguous()
    return v.view([-1] + list(v.shape[2:]))
