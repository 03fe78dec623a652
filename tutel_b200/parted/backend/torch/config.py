"""Code templates of the torch backend (reference: tutel/parted/backend/torch/config.py:9-75)."""
import os
import re
import sys


def get_input_definition(name, shape, stat_dim, dtype, is_param, device=None):
    return 'E.sharded_randn(%s, %s, dtype=torch.%s, requires_grad=True, is_param=%s, device=%s)' % (list(shape), stat_dim, dtype, is_param, device)


def get_execute_cmd(group_size, glob_size, device_type, program_path):
    env = 'PYTHONWARNINGS=ignore OMP_NUM_THREADS=1'
    if glob_size == 1:
        return '%s %s %s' % (env, sys.executable, program_path)
    hosts = os.environ.get('HOSTS', 'localhost').split(',')
    assert glob_size % len(hosts) == 0, 'Cannot evenly launch %d instances on %d hosts.' % (glob_size, len(hosts))
    local = glob_size // len(hosts)
    if len(hosts) == 1:   # single node: torchrun is enough (no mpiexec needed)
        port = 29000 + (os.getpid() % 2000)
        return '%s %s -m torch.distributed.run --nnodes=1 --nproc-per-node=%d --master-addr 127.0.0.1 --master-port %d %s' % (
            env, sys.executable, local, port, program_path)
    return 'mpiexec --allow-run-as-root -host %s -x MASTER_ADDR=%s -x LOCAL_SIZE=%d %s -m tutel_b200.launcher.run %s %s' % (
        ','.join(hosts), hosts[0], local, sys.executable, sys.executable, program_path)


def link(name, input_dim, output_dim, is_param=False, output_shape=None):
    """Code that takes `name` from sharding state `input_dim` to `output_dim` (None = partial-sum / gradient hook)."""
    grp = 'E.parallel_env.model_group'
    if input_dim is None:
        return 'C.allreduce_forward(%s, group=%s)' % (name, grp) if output_dim == -1 else 'C.reduce_scatter(%s, %s, %s)' % (name, output_dim, grp)
    if output_dim is None:
        return 'E.warp_bwd_allreduce(%s, %s)' % (name, is_param)
    if input_dim == -2:
        return 'C.zero_gather(%s, %s, %s)' % (name, list(output_shape), grp)
    if input_dim == -1:
        return 'C.spatial_split(%s, %s, %s)' % (name, output_dim, grp)
    if output_dim == -1:
        return 'C.all_gather(%s, %s, group=%s)' % (name, input_dim, grp)
    return 'C.all_to_all(%s, %s, %s, %s)' % (name, input_dim, output_dim, grp)


def generate_framework_code(device_type, group_size, group_count, run_mode, compute_name, headers, input_list, param_list, graph_prog):
    head = ('\n'.join(headers).strip() + '\n') if headers else ''
    body = '\n    '.join(graph_prog)
    for pname, _ in param_list:
        body = re.sub(r'\b%s\b' % re.escape(pname), 'self.%s' % pname, body)
    args = ', '.join(n for n, _ in input_list)
    inputs = '\n    '.join('inputs["%s"] = %s' % (n, c) for n, c in input_list) or 'pass'
    params = '\n    '.join('self.register_parameter(name="%s", param=%s)' % (n, c) for n, c in param_list) or 'pass'
    return '''import torch

from tutel_b200 import net as C
from tutel_b200.parted.backend.torch import executor as E

%s
class DistModel(torch.nn.Module):
  compute_name = '%s'

  def __init__(self):
    super().__init__()
    %s

  def forward(self, %s):
    %s
    return %s

  @staticmethod
  def synthetic_inputs():
    inputs = dict()
    %s
    return inputs


if __name__ == '__main__':
  E.init_session(group_size=%d, group_count=%d, device_type='%s')
  E.model_executor(DistModel, is_training=%s)
''' % (head, compute_name, params, args, body, compute_name, inputs, group_size, group_count, device_type, run_mode == 'train')
