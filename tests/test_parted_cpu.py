"""Auto-SPMD tool: layout parser, data-parallel config, code generation and a measured search on 1 CPU process."""
import os
import subprocess
import sys

from helpers import ROOT

SCRIPT = r'''
import os, sys, json
os.environ['DEVICE'] = 'cpu'
from tutel_b200.parted import spmdx
from tutel_b200.parted.layout import Layout
L = Layout('mn += mk, kn')
assert L.reduce_type == '+' and L.reduce_axes() == ['k'] and L.sources_of_output_dim(0) == ({0: 0, 1: -1}, 1)
assert Layout('(ab)c = abc').infer_shape([[2, 3, 4]]) == [6, 4]
spmdx.init('torch')
def Input(shape): return spmdx.Tensor(shape, 'float32', is_param=False)
def Param(shape): return spmdx.Tensor(shape, 'float32', is_param=True)
def Matmul(x, w): return spmdx.Custom('mn += mk, kn', f'torch.matmul({x}, {w})')
def Relu(x): return spmdx.Custom('mn = mn', f'torch.relu({x})')
y = Matmul(Relu(Matmul(Input([64, 32]), Param([32, 128]))), Param([128, 32]))
assert y.shape == [64, 32] and y.flops == 2 * 64 * 32 * 128
dp = y.get_data_parallel_config(total_nodes=2, spmd_nodes=2, device_type='cpu')
prog = y.compile(dp)          # 2-way data parallel program (generated, not run here)
assert 'class DistModel' in prog.code and 'torch.matmul' in prog.code and 'E.warp_bwd_allreduce' in prog.code
tp = {k: list(v) for k, v in dp.config['b'].items()}
names = sorted(tp)
# tensor parallel over the hidden dim: first matmul split on n (BAR), relu follows, second matmul contracts it (FAR)
relu = y.inputs[0]; mm0 = relu.inputs[0]; x, w0 = mm0.inputs; w1 = y.inputs[1]
tp.update({x.name: [-1, ''], w0.name: [1, ''], w1.name: [0, ''], mm0.name: [1, 'BAR:0'], relu.name: [1, 'BAR:0'], y.name: [-1, 'FAR:0']})
code = y.compile(tp, total_nodes=2, spmd_nodes=2, device_type='cpu', run_mode='train').code
assert 'C.allreduce_forward(%s' % y.name in code, code
cfg = y.autotune(total_nodes=1, spmd_nodes=1, device_type='cpu', config_file=sys.argv[1])
assert os.path.exists(sys.argv[1]) and cfg.config['v'] == '0.1' and y.name in cfg.config['b']
res = y.compile(cfg).execute()
assert res.get('step_time', 0) > 0, res
# the generated 2-way tensor-parallel and 2-way data-parallel programs really run (2 Gloo ranks) and compute the same
# first output element as the single-process program
res_tp = y.compile(tp, total_nodes=2, spmd_nodes=2, device_type='cpu', run_mode='train').execute()
res_dp = y.compile(dp).execute()
assert res_tp.get('step_time', 0) > 0 and res_dp.get('step_time', 0) > 0, (res_tp, res_dp)
assert abs(res_tp['digest'] - res['digest']) <= 1e-6 * max(1.0, abs(res['digest'])), (res_tp, res)
assert abs(res_dp['digest'] - res['digest']) <= 1e-6 * max(1.0, abs(res['digest'])), (res_dp, res)
# sharded code generation (not executed): 2-way tensor parallel hidden dim
cfg2 = {n: v for n, v in cfg.config['b'].items()}
print('PARTED_OK', json.dumps(cfg.config['b']))
'''


def test_parted_end_to_end(tmp_path):
    env = dict(os.environ)
    env['PYTHONPATH'] = ROOT + os.pathsep + env.get('PYTHONPATH', '')
    p = subprocess.run([sys.executable, '-c', SCRIPT, str(tmp_path / 'cfg.json')], capture_output=True, text=True,
                       timeout=600, env=env, cwd=ROOT)
    assert 'PARTED_OK' in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
