"""System utilities, compat shim, launcher argument plumbing, activation classification."""
import os

import pytest
import subprocess
import sys

import torch
import torch.nn.functional as F

from helpers import ROOT


def test_system_utils(tmp_path):
    from tutel_b200 import system
    assert system.apply_rank_size_from_pattern(str(tmp_path / 'a/{rank}-of-{size}.ckpt'), 3, 8).endswith('a/3-of-8.ckpt')
    assert os.path.isdir(tmp_path / 'a')
    t = torch.arange(6.).view(2, 3)
    system.save(t, str(tmp_path / 't.npy'))
    assert torch.equal(system.load(str(tmp_path / 't.npy')), t)
    c = system.cache()
    c.reset(); c.set('x', 1); c.set('y', 2)
    assert c.get('x') == 1 and sorted(c.get()) == [1, 2]
    assert system.record_time(is_cuda=False) > 0
    system.init_affinity_at_program_beginning()          # must not raise


def test_compat_shim_runs_reference_style_code():
    code = r'''
import tutel_b200.compat as compat
compat.install_as_tutel()
import torch, torch.nn.functional as F
from tutel import moe as tutel_moe, net, system, jit
from tutel.impls.fast_dispatch import extract_critical, fast_encode
from tutel.impls import communicate as C
from tutel.experts.ffn import ExpertModule
from tutel.gates.top import Gate
env = system.init_data_model_parallel(backend='gloo')
layer = tutel_moe.moe_layer(gate_type={'type': 'top', 'k': 2}, model_dim=8,
                            experts={'type': 'ffn', 'count_per_node': 2, 'hidden_size_per_expert': 8, 'activation_fn': lambda x: F.relu(x)})
y = layer(torch.randn(4, 8))
assert y.shape == (4, 8) and C.get_world_size() == 1 and net.simple_all_reduce(y) is y
print('COMPAT_OK')
'''
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert 'COMPAT_OK' in p.stdout, p.stdout + p.stderr


def test_activation_classification():
    from tutel_b200.ops.gemm import classify_activation
    assert classify_activation(None) == 'relu' and classify_activation(F.relu) == 'relu'
    assert classify_activation(lambda x: F.relu(x)) == 'relu'
    assert classify_activation(lambda x: F.gelu(x)) == 'gelu' and classify_activation(F.silu) == 'silu'
    assert classify_activation(lambda x: F.silu(x)) == 'silu' and classify_activation(torch.tanh) is None
    drop = torch.nn.Dropout(0.5)
    assert classify_activation(lambda x: drop(F.relu(x))) is None      # stochastic -> never fused
    # clamped variants agree with ReLU on small inputs only: they must not be taken for ReLU
    assert classify_activation(F.relu6) is None and classify_activation(lambda x: torch.clamp(x, 0, 5)) is None
    assert classify_activation(lambda x: F.hardtanh(x, 0.0, 1000.0)) is None


def test_launcher_builds_torchrun_command(monkeypatch):
    from tutel_b200.launcher import run
    seen = {}
    monkeypatch.setattr(os, 'execvpe', lambda f, a, e: seen.update(cmd=a, env=dict(e)))
    monkeypatch.setattr(sys, 'argv', ['run', '-m', 'my.prog', '--flag'])
    monkeypatch.setenv('OMPI_COMM_WORLD_SIZE', '2')
    monkeypatch.setenv('OMPI_COMM_WORLD_RANK', '1')
    monkeypatch.setenv('LOCAL_SIZE', '4')
    run.main()
    cmd = seen['cmd']
    assert '--nproc_per_node=4' in cmd and '--nnodes=2' in cmd and '--node_rank=1' in cmd
    assert cmd[-3:] == ['-m', 'my.prog', '--flag'] and 'tutel_b200.launcher.execl' in cmd
    from tutel_b200.launcher import execl
    monkeypatch.setattr(sys, 'argv', ['execl', '-m', 'my.prog'])
    monkeypatch.setenv('TUTEL_CUDA_SANDBOX', '2')
    monkeypatch.setenv('LOCAL_RANK', '3')
    execl.main()
    assert seen['env']['CUDA_VISIBLE_DEVICES'] == '3' and seen['cmd'][-2:] == ['-m', 'my.prog']
    assert seen['cmd'][0] == sys.executable
    # reference-style `launcher.run python3 train.py ..`: the command is exec'd as given, no interpreter is prepended
    monkeypatch.setenv('NUMA_TYPE', '0')
    monkeypatch.setattr(sys, 'argv', ['execl', 'python3', 'train.py', '--x'])
    execl.main()
    assert seen['cmd'] == ['python3', 'train.py', '--x']


def test_fairseq_style_conversion(monkeypatch):
    """examples/fairseq_moe: in-place conversion of fairseq-style transformer layers + aux-loss hook."""
    from tutel_b200 import system
    from tutel_b200.examples.fairseq_moe import add_moe_aux_loss, convert_transformer_layers, zero_overflow_grads

    class Layer(torch.nn.Module):          # fairseq's TransformerDecoderLayerBase attribute layout
        def __init__(self, d=16, h=32):
            super().__init__()
            self.embed_dim, self.quant_noise = d, 0
            self.fc1, self.fc2 = torch.nn.Linear(d, h), torch.nn.Linear(h, d)
            self.activation_fn = F.relu
            self.activation_dropout_module = torch.nn.Dropout(0.0)
            self.ffn_layernorm = torch.nn.LayerNorm(h)

        def forward(self, x):
            residual = x
            x = self.activation_fn(self.fc1(x))
            x = self.activation_dropout_module(x)
            if self.ffn_layernorm is not None:
                x = self.ffn_layernorm(x)
            return residual + self.fc2(x)

    torch.manual_seed(0)
    model = torch.nn.Sequential(*[Layer() for _ in range(4)])
    monkeypatch.setenv('MOE', '2')
    assert convert_transformer_layers(model) == 2
    assert type(model[1].fc1).__name__ == 'MoEFeedForward' and isinstance(model[0].fc1, torch.nn.Linear)
    assert all(getattr(p, 'expert', False) for n, p in model[1].fc1.moe_ffn.named_parameters() if 'experts' in n)
    system.cache().reset()
    x = torch.randn(3, 5, 16)
    y = model(x)
    assert y.shape == x.shape
    records = system.cache().get()
    assert len(records) == 2 and records[0][0] == 15
    loss = add_moe_aux_loss(y.pow(2).mean(), l_aux_wt=0.01)
    assert system.cache().get() == []
    loss.backward()
    assert model[1].fc1.moe_ffn.gates[0].wg.weight.grad is not None
    assert model[1].fc1.ffn_layernorm.weight.grad is not None       # the moved layer-norm still trains
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.tensor([1.0, float('inf'), -float('inf')])
    zero_overflow_grads([p], enabled=True)
    assert p.grad.tolist() == [1.0, 0.0, 0.0]


def test_block_pool_first_fit_coalescing_and_no_overlap():
    """csrc/symm_heap.cpp BlockPool: the allocator behind the zero-copy receive buffers of the P2P collectives."""
    import random
    from tutel_b200.ops import backend
    ext = backend.ext()
    if ext is None or not hasattr(ext, 'BlockPool'):
        pytest.skip('native extension not built')
    pool = ext.BlockPool()
    base, size = 4096, 1 << 20
    pool.reset(base, size)
    assert pool.alloc(size + 1) == -1 and pool.free_bytes() == size
    a = pool.alloc(1)                       # rounded up to the 256-byte granule
    b = pool.alloc(1000)
    assert a == base and b == base + 256 and pool.live_blocks() == 2
    pool.free(a)
    assert pool.alloc(300) == base + 256 + 1024      # first fit skips the 256-byte hole
    assert pool.alloc(256) == base                   # ... which an exact fit re-uses
    pool.free(12345)                                 # unknown offsets are ignored
    pool.reset(base, size)
    rng = random.Random(0)
    live = {}
    for step in range(4000):
        if live and (rng.random() < 0.45 or pool.largest_free_block() < 256):
            off = rng.choice(list(live))
            pool.free(off)
            del live[off]
        else:
            n = rng.choice([1, 100, 256, 257, 4096, 70000])
            off = pool.alloc(n)
            if off < 0:
                assert pool.largest_free_block() < (n + 255) // 256 * 256
                continue
            length = (n + 255) // 256 * 256
            assert off % 256 == 0 and base <= off and off + length <= base + size
            for o, l in live.items():
                assert off + length <= o or o + l <= off, 'blocks overlap'
            live[off] = length
        assert pool.free_bytes() == size - sum(live.values()) and pool.live_blocks() == len(live)
    for off in list(live):
        pool.free(off)
    assert pool.free_bytes() == size and pool.largest_free_block() == size      # everything coalesced again


def test_jit_launch_annotation_scanner():
    """csrc/jit_nvrtc.cpp: `// [thread_extent] blockIdx.x = N` annotations + entry-point detection (reference:
    tutel/custom/custom_kernel.cpp:174-218 does this with sscanf / strstr)."""
    from tutel_b200.ops import backend
    ext = backend.ext()
    if ext is None or not hasattr(ext, 'jit_parse'):
        pytest.skip('native extension not built')
    src = r'''
      #include <cuda_fp16.h>
      static __device__ float helper(float v) { return v * 2; }
      extern "C" __global__ __launch_bounds__(64) void   my_kernel_7(float* __restrict__ x, int n) {
        // [thread_extent] blockIdx.x = 512
        // [thread_extent]   threadIdx.x=64
        // [thread_extent] blockIdx.y = 3
        // [thread_extent] bogusIdx.x = 9
        // [thread_extent] threadIdx.z =
        x[0] = helper(x[0]);
      }'''
    entry, grid, block = ext.jit_parse(src)
    assert entry == 'my_kernel_7' and grid == [512, 3, 1] and block == [64, 1, 1]
    entry, grid, block = ext.jit_parse('__global__ void k(int*a){}')
    assert entry == 'k' and grid == [1, 1, 1] and block == [1, 1, 1]
    with pytest.raises(RuntimeError):
        ext.jit_parse('void not_a_kernel(int* a) {}')
    with pytest.raises(RuntimeError):
        ext.jit_parse('// [thread_extent] blockIdx.x = 4')


def test_fused_ring_spills_one_transaction_ahead():
    """parallel/fused.py: transactions take buffer sets round robin; a set that still carries a lease (a forward whose
    backward is pending) is spilled when it becomes the NEXT set - one transaction before it is re-used - and a geometry's
    ring is keyed by the power-of-two row quantum, not by the exact capacity."""
    import weakref
    from tutel_b200.parallel import fused

    class FakeTransport:
        def __init__(self):
            self.bump, self.ctrl = 0, 0

        def can_alloc(self, n):
            return True

        def alloc(self, name, n):
            self.bump += n
            return self.bump - n

        def ctrl_alloc(self, name, n):
            self.ctrl += n
            return self.ctrl - n

    class FakeLease:
        def __init__(self, bufs, log, tag):
            self.bufs, self.log, self.tag = bufs, log, tag

        def spill(self, stream):
            self.log.append(self.tag)
            self.bufs.holder = None

    geo = fused._Geometry(W=2, rank=0, E=4, El=2, C=300, k=2, M=64, H=128, Mo=64, es=2)
    assert geo.rows_quantum() == 512 and geo.key() == fused._Geometry(2, 0, 4, 2, 500, 2, 64, 128, 64, 2).key()
    assert geo.key() != fused._Geometry(2, 0, 4, 2, 513, 2, 64, 128, 64, 2).key()
    ring = fused._Ring(FakeTransport(), geo, 1)
    assert len(ring.sets) == 3 and ring.usable()
    log, leases = [], []

    def forward(tag):
        s = ring.next(None)
        assert s.holder is None
        lease = FakeLease(s, log, tag)
        s.holder = weakref.ref(lease)
        leases.append(lease)
        return s

    def backward(lease):
        s = ring.next(None)
        assert s.holder is None
        if lease.bufs.holder is not None and lease.bufs.holder() is lease:
            lease.bufs.holder = None
        return s

    # one layer: forward on set 0, backward on set 1, next forward on set 2, ... - nothing is ever spilled
    for step in range(4):
        forward('L0s%d' % step)
        backward(leases[-1])
    assert log == []
    # four layers deep: the ring wraps while leases are alive -> each is spilled exactly once, before its set is handed out
    log.clear(), leases.clear()
    for layer in range(4):
        forward('f%d' % layer)
    assert log == ['f0', 'f1']               # set of f0 became "next" during f2, set of f1 during f3
    for lease in reversed(leases):
        backward(lease)
    assert sorted(log) == ['f0', 'f1', 'f2', 'f3'][:len(log)] and len(log) == len(set(log))
