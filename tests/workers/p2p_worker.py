"""Worker for multi-GPU tests: P2P collectives vs torch.distributed (NCCL), fused MoE path vs NCCL path."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tutel_b200 import moe, net, system  # noqa: E402
from tutel_b200.parallel import p2p  # noqa: E402


_FAILED = []


def check(name, ok):
    """Record the verdict and keep going: a rank that stopped here would leave its peers waiting in the next collective
    (main() turns recorded failures into a non-zero exit code on every rank)."""
    r = dist.get_rank()
    print('[rank %d] %s: %s' % (r, name, 'OK' if ok else 'FAIL'), flush=True)
    if not ok:
        _FAILED.append(name)


def test_collectives(env):
    W, r, dev = env.global_size, env.global_rank, env.local_device
    t = p2p.transport_for(None)
    check('transport', t is not None)
    torch.manual_seed(100 + r)
    for n in (W * 4, W * 1000, W * 65536 + W * 3):
        x = torch.randn(n, device=dev)
        ref = torch.empty_like(x)
        dist.all_to_all_single(ref, x)
        check('all_to_all n=%d' % n, torch.equal(net.simple_all_to_all(x), ref))
    x = torch.randn(W, 3, 5, device=dev).bfloat16()
    ref = torch.empty_like(x)
    dist.all_to_all_single(ref, x)
    check('all_to_all bf16', torch.equal(net.simple_all_to_all(x), ref))
    y = torch.randn(7, 3, device=dev)
    ref = torch.empty(W * 7, 3, device=dev)
    dist.all_gather_into_tensor(ref, y)
    check('all_gather', torch.equal(net.simple_all_gather(y), ref))
    z = torch.randn(W * 6, 4, device=dev)
    ref = z.clone()
    dist.all_reduce(ref)
    check('all_reduce', torch.allclose(net.simple_all_reduce(z), ref, atol=1e-5))
    ref_rs = ref.view(W, -1)[r].view(6, 4)
    check('reduce_scatter', torch.allclose(net.simple_reduce_scatter(z), ref_rs, atol=1e-5))
    zm = torch.randn(33, device=dev)
    refm = zm.clone()
    dist.all_reduce(refm, op=dist.ReduceOp.MAX)
    check('all_reduce max', torch.equal(net.simple_all_reduce(zm, op=dist.ReduceOp.MAX), refm))
    # one-shot all-reduce kernel: dtypes, ops, sizes around its 256 KiB limit, repeated calls (buffer parity)
    for dt, n, op in ((torch.int32, 1, dist.ReduceOp.MAX), (torch.int64, 5, dist.ReduceOp.SUM), (torch.bfloat16, 32768, dist.ReduceOp.SUM),
                      (torch.float16, 1001, dist.ReduceOp.MAX), (torch.float32, 65536, dist.ReduceOp.SUM),
                      (torch.float32, 65537, dist.ReduceOp.SUM), (torch.float32, 3, dist.ReduceOp.SUM)):
        for rep in range(3):
            v = (torch.randn(n, device=dev) * 8).to(dt) if dt.is_floating_point else torch.randint(-99, 99, [n], device=dev, dtype=dt) + r
            refv = v.clone()
            dist.all_reduce(refv, op=op)
            got = net.simple_all_reduce(v, op=op)
            tol = 0 if not dt.is_floating_point or op == dist.ReduceOp.MAX else (1e-5 if dt == torch.float32 else 0.13)
            check('all_reduce %s n=%d op=%s #%d' % (str(dt).split('.')[-1], n, 'max' if op == dist.ReduceOp.MAX else 'sum', rep),
                  torch.allclose(got.float(), refv.float(), atol=tol * 8, rtol=tol) and got.data_ptr() != v.data_ptr())
    w = torch.randn(4, 4096, device=dev).bfloat16()
    outs = [torch.empty_like(w) for _ in range(W)]
    mine = net.simple_all_reduce(w)
    dist.all_gather(outs, mine)
    check('all_reduce bit-identical on every rank', all(torch.equal(o, mine) for o in outs))
    # ragged
    counts = [(r + p) % 3 + 1 for p in range(W)]
    data = torch.arange(sum(counts), device=dev, dtype=torch.float32) + 100 * r
    (out,), sizes = net.batch_all_to_all_v([data], counts)
    want = []
    for s in range(W):
        c = [(s + p) % 3 + 1 for p in range(W)]
        full = torch.arange(sum(c), dtype=torch.float32) + 100 * s
        off = sum(c[:r])
        want.append(full[off:off + c[r]])
    check('all_to_all_v', torch.equal(out.cpu(), torch.cat(want)) and sizes.tolist() == [len(w) for w in want])
    mine = torch.full([r + 2], float(r), device=dev)
    (g,), gs = net.batch_all_gather_v([mine])
    check('all_gather_v', torch.equal(g.cpu(), torch.cat([torch.full([s + 2], float(s)) for s in range(W)])))
    # autograd all_to_all round trip
    a = torch.randn(W * 2, 3, 4, device=dev, requires_grad=True)
    b = net.all_to_all(net.all_to_all(a, 1, 0), 0, 1)
    b.sum().backward()
    check('all_to_all autograd', torch.equal(b, a) and torch.equal(a.grad, torch.ones_like(a)))


def run_layer(env, fused, dtype, nle, steps=3, overlap=1, model_dim=256, hidden=512, tokens=512, expert='ffn', act=None,
              is_postscore=True, output_dim=None, parallel_type='adaptive:1', same_experts=False, fp8=False):
    os.environ['TUTEL_B200_FUSED'] = '1' if fused else '0'
    r, dev = env.global_rank, env.local_device
    torch.manual_seed(7)
    experts = {'type': expert, 'num_experts_per_device': nle, 'hidden_size_per_expert': hidden,
               'activation_fn': act or (lambda t: F.relu(t))}
    if output_dim is not None:
        experts['output_dim'] = output_dim
    if fp8:
        experts['fp8'] = fp8          # True / 'row': row scales (fused engine); 'mx': MX block scales (unfused path)
    layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 1.5}, model_dim=model_dim, experts=experts,
                          seeds=(1, 1 if same_experts else r + 1, 1), a2a_ffn_overlap_degree=overlap, is_postscore=is_postscore,
                          parallel_type=parallel_type).to(dev).to(dtype)
    opt = torch.optim.SGD(layer.parameters(), lr=1e-2)
    torch.manual_seed(50 + r)
    x = torch.randn(tokens, model_dim, device=dev).to(dtype)
    losses, grads = [], None
    for _ in range(steps):
        opt.zero_grad()
        xi = x.clone().requires_grad_(True)
        y = layer(xi)
        loss = y.float().pow(2).mean() + 0.01 * y.l_aux.float()
        loss.backward()
        losses.append(float(loss))
        grads = xi.grad.float().clone()
        opt.step()
    return losses, grads, y.float().detach()


def test_fused_vs_nccl(env):
    for dtype in (torch.bfloat16,):
        for nle in (1, 2):
            for overlap in (1, 2):
                a = run_layer(env, True, dtype, nle, overlap=overlap)
                b = run_layer(env, False, dtype, nle, overlap=1)
                ok = all(abs(u - v) <= 2e-2 * max(1.0, abs(v)) for u, v in zip(a[0], b[0]))
                ok = ok and torch.allclose(a[2], b[2], atol=3e-2, rtol=3e-2) and torch.allclose(a[1], b[1], atol=3e-2, rtol=5e-2)
                if not ok:
                    print('losses', a[0], b[0], (a[2] - b[2]).abs().max().item(), (a[1] - b[1]).abs().max().item(), flush=True)
                check('fused==nccl nle=%d d=%d' % (nle, overlap), ok)
    # every mode the engine covers: pre-score gating, other activations, output_dim, experts sharded over the GPUs
    W = env.global_size
    modes = [('prescore ffn', dict(expert='ffn', is_postscore=False)), ('prescore llama_ffn', dict(expert='llama_ffn', is_postscore=False)),
             ('gelu', dict(act=F.gelu)), ('silu', dict(act=lambda t: F.silu(t))), ('output_dim', dict(output_dim=128)),
             ('sharded r=1 (data)', dict(nle=-W, parallel_type='data', hidden=512, same_experts=True)),
             ('sharded r=W (model)', dict(nle=-W, parallel_type='model', hidden=512, same_experts=True)),
             ('sharded llama_ffn', dict(nle=-W, expert='llama_ffn', act=F.silu, same_experts=True))]
    for name, kw in modes:
        kw = dict(kw)
        nle = kw.pop('nle', 1)
        a = run_layer(env, True, torch.bfloat16, nle, **kw)
        b = run_layer(env, False, torch.bfloat16, nle, **kw)
        ok = all(abs(u - v) <= 2e-2 * max(1.0, abs(v)) for u, v in zip(a[0], b[0]))
        ok = ok and torch.allclose(a[2], b[2], atol=3e-2, rtol=3e-2) and torch.allclose(a[1], b[1], atol=3e-2, rtol=5e-2)
        if not ok:
            print(name, 'losses', a[0], b[0], (a[2] - b[2]).abs().max().item(), (a[1] - b[1]).abs().max().item(), flush=True)
        check('fused==generic %s' % name, ok)
    # gated (SwiGLU) experts through the same engine
    for nle, act in ((1, F.silu), (2, None)):
        a = run_layer(env, True, torch.bfloat16, nle, expert='llama_ffn', act=act)
        b = run_layer(env, False, torch.bfloat16, nle, expert='llama_ffn', act=act)
        ok = all(abs(u - v) <= 2e-2 * max(1.0, abs(v)) for u, v in zip(a[0], b[0]))
        ok = ok and torch.allclose(a[2], b[2], atol=3e-2, rtol=3e-2) and torch.allclose(a[1], b[1], atol=3e-2, rtol=5e-2)
        if not ok:
            print('llama losses', a[0], b[0], (a[2] - b[2]).abs().max().item(), (a[1] - b[1]).abs().max().item(), flush=True)
        check('fused==nccl llama_ffn nle=%d' % nle, ok)


def test_subgroups(env):
    """Model / data sub-groups get their own P2P transports (own arenas, exchanged among the members only)."""
    W, r, dev = env.global_size, env.global_rank, env.local_device
    g = net.create_groups_from_world(group_count=2)
    for name, grp in (('model', g.model_group), ('data', g.data_group)):
        n = dist.get_world_size(grp)
        t = p2p.transport_for(grp)
        check('sub-group transport (%s, %d ranks)' % (name, n), (t is not None) == (n > 1) and (t is None or t.world == n))
        if n == 1:
            continue
        torch.manual_seed(300 + r)
        x = torch.randn(n * 1000, device=dev)
        ref = torch.empty_like(x)
        dist.all_to_all_single(ref, x, group=grp)
        check('%s-group all_to_all' % name, torch.equal(net.simple_all_to_all(x, group=grp), ref))
        z = torch.randn(n * 6, 4, device=dev)
        refz = z.clone()
        dist.all_reduce(refz, group=grp)
        check('%s-group all_reduce' % name, torch.allclose(net.simple_all_reduce(z, group=grp), refz, atol=1e-5))
        y = torch.randn(5, 3, device=dev).bfloat16()
        refy = torch.empty(n * 5, 3, device=dev, dtype=torch.bfloat16)
        dist.all_gather_into_tensor(refy, y, group=grp)
        check('%s-group all_gather' % name, torch.equal(net.simple_all_gather(y, group=grp), refy))


def test_2dh(env):
    """Hierarchical all-to-all (LOCAL_SIZE=2 simulates 2 nodes x 2 GPUs): stride-copy kernel + two sub-group phases."""
    W, r, dev = env.global_size, env.global_rank, env.local_device
    torch.manual_seed(400 + r)
    x = torch.randn(W * 2, 6, 16, device=dev).bfloat16()
    flat = net.all_to_all(x, 1, 0, use_2dh=False)
    hier = net.all_to_all(x, 1, 0, use_2dh=True)
    check('2dh == flat (1,0)', torch.equal(flat, hier))
    check('2dh == flat (0,1)', torch.equal(net.all_to_all(flat, 0, 1, use_2dh=True), net.all_to_all(flat, 0, 1)))
    os.environ['TUTEL_B200_FUSED'] = '0'
    outs = []
    for use_2dh, d in ((False, 1), (True, 1), (True, 2)):
        torch.manual_seed(7)
        layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2}, model_dim=64, use_2dh=use_2dh, a2a_ffn_overlap_degree=d,
                              experts={'type': 'ffn', 'num_experts_per_device': 2, 'hidden_size_per_expert': 128,
                                       'activation_fn': lambda t: F.relu(t)}, seeds=(1, r + 1, 1)).to(dev)
        torch.manual_seed(50 + r)
        xi = torch.randn(256, 64, device=dev, requires_grad=True)
        y = layer(xi)
        y.pow(2).mean().backward()
        outs.append((y.detach(), xi.grad.clone(), layer.experts.batched_fc1_w.grad.clone()))
    for tag, o in (('2dh d=1', outs[1]), ('2dh d=2 (async phases)', outs[2])):
        check('layer %s == flat' % tag, all(torch.allclose(a, b, atol=1e-5, rtol=1e-4) for a, b in zip(o, outs[0])))
    os.environ['TUTEL_B200_FUSED'] = '1'


def test_fp8(env):
    """e4m3 experts inside the fused engine (rows quantised in the push, fp8 forward + data-gradient GEMMs, weights quantised
    once per step): same results as the unfused fp8 path, and a loss curve that tracks the bf16 run."""
    for expert, act in (('ffn', None), ('llama_ffn', F.silu)):
        a = run_layer(env, True, torch.bfloat16, 2, expert=expert, act=act, fp8=True, steps=6)
        b = run_layer(env, False, torch.bfloat16, 2, expert=expert, act=act, fp8=True, steps=6)
        c = run_layer(env, True, torch.bfloat16, 2, expert=expert, act=act, fp8=False, steps=6)
        ok = all(abs(u - v) <= 3e-2 * max(1.0, abs(v)) for u, v in zip(a[0], b[0]))
        ok = ok and torch.allclose(a[2], b[2], atol=6e-2, rtol=6e-2) and torch.allclose(a[1], b[1], atol=6e-2, rtol=1e-1)
        if not ok:
            print('fp8', expert, a[0], b[0], (a[2] - b[2]).abs().max().item(), (a[1] - b[1]).abs().max().item(), flush=True)
        check('fused fp8 == unfused fp8 (%s)' % expert, ok)
        drift = max(abs(u - v) / max(1e-6, abs(v)) for u, v in zip(a[0], c[0]))
        print('[rank %d] fp8 vs bf16 losses (%s): %s vs %s (max rel. diff %.4f)' % (env.global_rank, expert, a[0], c[0], drift), flush=True)
        check('fp8 loss curve within 5%% of bf16 (%s)' % expert, drift < 5e-2)
    # MX block-scaled experts (fp8='mx'): the fused engine declines them, the generic exchange feeds the MX GEMMs
    m = run_layer(env, True, torch.bfloat16, 2, model_dim=256, hidden=512, fp8='mx', steps=6)
    c = run_layer(env, True, torch.bfloat16, 2, model_dim=256, hidden=512, fp8=False, steps=6)
    drift = max(abs(u - v) / max(1e-6, abs(v)) for u, v in zip(m[0], c[0]))
    print('[rank %d] mx vs bf16 losses: %s vs %s (max rel. diff %.4f)' % (env.global_rank, m[0], c[0], drift), flush=True)
    check('mx loss curve within 5% of bf16', drift < 5e-2)
    check('mx output close to bf16', ((m[2] - c[2]).norm() / c[2].norm()).item() < 0.1)


def test_deep_stack(env):
    """Six MoE layers share ONE ring of three buffer sets: every layer's lease is spilled when the ring wraps, and the
    result must match the generic path (which keeps everything in ordinary tensors)."""
    r, dev = env.global_rank, env.local_device
    from tutel_b200.parallel import fused as fused_mod
    outs = []
    for fused in (True, False):
        os.environ['TUTEL_B200_FUSED'] = '1' if fused else '0'
        torch.manual_seed(7)
        layers = torch.nn.ModuleList([
            moe.moe_layer(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 1.25}, model_dim=256,
                          experts={'type': 'ffn', 'num_experts_per_device': 2, 'hidden_size_per_expert': 512,
                                   'activation_fn': lambda t: F.relu(t)}, seeds=(1, r + 1, 1)) for _ in range(6)]).to(dev).bfloat16()
        torch.manual_seed(50 + r)
        x = torch.randn(512, 256, device=dev).bfloat16().requires_grad_(True)
        h = x
        for layer in layers:
            h = h + layer(h)
        h.float().pow(2).mean().backward()
        outs.append((h.float().detach(), x.grad.float().clone(), layers[0].experts.batched_fc1_w.grad.float().clone(),
                     layers[5].experts.batched_fc2_w.grad.float().clone()))
        if fused:
            t = p2p.transport_for(None)
            rings = [rg for rg in fused_mod._engine(t).rings.values() if rg is not None]
            check('one shared ring of <= 3 sets for 6 layers', len(rings) >= 1 and all(len(rg.sets) <= 3 for rg in rings))
    os.environ['TUTEL_B200_FUSED'] = '1'

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()
    errs = [rel(a, b) for a, b in zip(outs[0], outs[1])]
    print('[rank %d] deep stack rel errors %s' % (r, errs), flush=True)
    check('6-layer stack fused == generic', max(errs) < 3e-2)


def _oracle_pair(env, nle, expert, tokens, model_dim, hidden, k=2, act=None):
    """Fused bf16 engine vs a plain-PyTorch fp32 oracle (NCCL all-to-all + torch.matmul experts) on IDENTICAL,
    bf16-representable weights / inputs and an fp32 gate (so both runs take the same routing decisions)."""
    r, dev = env.global_rank, env.local_device
    res = []
    for fused in (True, False):
        os.environ['TUTEL_B200_FUSED'] = '1' if fused else '0'
        os.environ['TUTEL_B200_COMM'] = 'p2p' if fused else 'nccl'
        torch.manual_seed(7)
        layer = moe.moe_layer(gate_type={'type': 'top', 'k': k, 'fp32_gate': True, 'capacity_factor': 1.25}, model_dim=model_dim,
                              experts={'type': expert, 'num_experts_per_device': nle, 'hidden_size_per_expert': hidden,
                                       'activation_fn': act or (lambda t: F.relu(t))}, seeds=(1, r + 1, 1)).to(dev)
        with torch.no_grad():
            for prm in layer.parameters():
                prm.copy_(prm.bfloat16().float())
        if fused:
            layer = layer.bfloat16()
        torch.manual_seed(50 + r)
        x = torch.randn(tokens, model_dim, device=dev).bfloat16()
        xi = (x if fused else x.float()).clone().requires_grad_(True)
        y = layer(xi)
        dy = torch.randn(tokens, model_dim, device=dev, generator=torch.Generator(device=dev).manual_seed(9 + r)).bfloat16()
        (y.float() * dy.float()).sum().backward()
        ex = {n: p_.grad.float().clone() for n, p_ in layer.experts.named_parameters()}
        res.append((y.float().detach(), xi.grad.float().clone(), ex, layer.gates[0].wg.weight.grad.float().clone()))
        del layer
    os.environ['TUTEL_B200_FUSED'], os.environ['TUTEL_B200_COMM'] = '1', 'p2p'

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-12)).item()
    errs = {'y': rel(res[0][0], res[1][0]), 'dx': rel(res[0][1], res[1][1]), 'dwg': rel(res[0][3], res[1][3])}
    for n in res[0][2]:
        errs['d' + n] = rel(res[0][2][n], res[1][2][n])
    return errs


def test_oracle(env):
    big = os.environ.get('TUTEL_B200_TEST_FLAGSHIP', '1') == '1'
    cases = [('ffn', 2, 512, 256, 512), ('llama_ffn', 1, 512, 256, 512)]
    if big:
        cases.append(('ffn', 8 // env.global_size if env.global_size <= 8 else 1, 8192, 4096, 14336))     # flagship shape
    for expert, nle, tokens, M, H in cases:
        errs = _oracle_pair(env, nle, expert, tokens, M, H, act=F.silu if expert == 'llama_ffn' else None)
        worst = max(errs.values())
        print('[rank %d] oracle %s M=%d H=%d rel errors: %s' % (env.global_rank, expert, M, H,
                                                                 ' '.join('%s=%.4f' % kv for kv in sorted(errs.items()))), flush=True)
        check('fused bf16 == fp32 torch oracle (%s, M=%d, H=%d)' % (expert, M, H), worst < 2e-2)


def test_parallel_equivalence(env):
    """fp32 on GPUs: a sharded expert gives the same losses under data and model parallelism, and for overlap 1 and 2
    (the reference's GPU equivalence checks, tests/test_tutel.py:154-176)."""
    W, r, dev = env.global_size, env.global_rank, env.local_device
    runs = {}
    for ptype, d in (('data', 1), ('model', 1), ('model', 2)):
        torch.manual_seed(3)
        layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2}, model_dim=128, parallel_type=ptype, a2a_ffn_overlap_degree=d,
                              experts={'type': 'ffn', 'num_experts_per_device': -W, 'hidden_size_per_expert': 64 * W,
                                       'activation_fn': lambda t: F.relu(t)}, seeds=(1, 1, 1)).to(dev)
        opt = torch.optim.SGD(layer.parameters(), lr=1e-2)
        torch.manual_seed(11)
        x = torch.randn(256, 128, device=dev)
        ls = []
        for _ in range(4):
            opt.zero_grad()
            y = layer(x)
            loss = y.pow(2).mean()
            loss.backward()
            if W > 1:
                for prm in layer.gates.parameters():
                    prm.grad = net.simple_all_reduce(prm.grad) / W
            opt.step()
            ls.append(float(loss))
        runs[(ptype, d)] = ls
    base = runs[('data', 1)]
    for key, ls in runs.items():
        check('sharded expert %s d=%d == data-parallel losses' % key, all(abs(a - b) <= 1e-5 * max(1, abs(b)) for a, b in zip(ls, base)))


def main():
    env = system.init_data_model_parallel(backend='nccl')
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'coll'):
        test_collectives(env)
    if which in ('all', 'fused'):
        test_fused_vs_nccl(env)
    if which in ('all', 'oracle'):
        test_oracle(env)
    if which in ('all', 'deep'):
        test_deep_stack(env)
    if which in ('all', 'fp8'):
        test_fp8(env)
    if which in ('all', 'equiv'):
        test_parallel_equivalence(env)
    if which in ('sub',) or (which == 'all' and env.global_size >= 4):
        test_subgroups(env)
    if which in ('2dh',) or (which == 'all' and env.global_size >= 4 and os.environ.get('LOCAL_SIZE') == '2'):
        test_2dh(env)
    if which == 'fault':
        # launched with TUTEL_B200_FAULT=skip_push:rank=1:call=2 - the second all-to-all must end in a diagnosed timeout
        a = torch.ones(1 << 16, device=env.local_device)
        net.simple_all_to_all(a)
        torch.cuda.synchronize()
        print('FIRST_OK', flush=True)
        try:
            net.simple_all_to_all(a)
            torch.cuda.synchronize()     # raises on the surviving ranks (device trap after the bounded spin)
            import time
            time.sleep(20)               # the rank that skipped its push: wait for the launcher to tear the job down
        finally:
            sys.stdout.flush()
            os._exit(17)                 # the CUDA context is gone: no destructors, no NCCL teardown
    bad = torch.tensor([len(_FAILED)], device=env.local_device)
    dist.all_reduce(bad)
    if int(bad.item()) > 0:
        print('[rank %d] FAILED checks: %s' % (env.global_rank, _FAILED), flush=True)
        raise SystemExit(3)
    if env.global_rank == 0:
        print('WORKER_OK', flush=True)


if __name__ == '__main__':
    main()
