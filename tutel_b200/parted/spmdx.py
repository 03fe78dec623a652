"""Graph DSL + compiler front-end of the auto-SPMD tool (API parity with tutel/parted/spmdx.py).

    from tutel_b200.parted import spmdx
    spmdx.init('torch')
    def Input(shape): return spmdx.Tensor(shape, 'float32', is_param=False)
    def Param(shape): return spmdx.Tensor(shape, 'float32', is_param=True)
    def Matmul(x, w): return spmdx.Custom('mn += mk, kn', f'torch.matmul({x}, {w})')
    def Relu(x):      return spmdx.Custom('mn = mn', f'torch.relu({x})')
    y = Matmul(Relu(Matmul(Input([64, 32]), Param([32, 128]))), Param([128, 32]))
    cfg = y.autotune(total_nodes=2, spmd_nodes=2, device_type='cpu')
    print(y.compile(cfg).code)

The op name is the name of the *calling function* (must start with an upper-case letter); ``str(node)`` renders as
``@@name@@`` so that f-strings build the forward expression.
"""
import importlib
import inspect
import json
import logging
import os
import re
import tempfile

from . import patterns, solver  # noqa: F401
from .layout import Layout

session = None


class _Session:
    def __init__(self, backend_name):
        self.backend = importlib.import_module('.backend.%s.config' % backend_name, __package__)
        self.is_strict_fmt = int(os.environ.get('STRICT_FMT', 0)) > 0
        self.ptype = os.environ.get('PTYPE', '')
        self.custom_dict = {}
        manual = json.loads(os.environ.get('CONFIG', '') or '{}')
        self.manual_config = {k: (v if isinstance(v, int) else v[0]) for k, v in manual.items()}
        self.counters = {'op': 0, 'id': 0, 'builtin': 0}


def init(backend_name):
    global session
    if session is not None:
        raise Exception('Function `init()` can be only invoked once.')
    if not re.match('^[a-zA-Z0-9]+$', backend_name):
        raise Exception('Only letters and digits are allowed for backend_name, get: %s' % backend_name)
    session = _Session(backend_name)
    try:
        return importlib.import_module('.backend.%s' % backend_name, __package__)
    except Exception:  # noqa
        return None


def reset():
    """Forget the current session (test helper; the reference has no equivalent)."""
    global session
    session = None


def new_dependency(header_content, depends=()):
    depends = list(depends) if isinstance(depends, (list, tuple)) else [depends]
    return {'data': header_content.strip() + '\n', 'depends': depends}


def product(values):
    n = 1
    for v in values:
        n *= int(v)
    return n


class Program:
    def __init__(self, code, kwargs):
        self.code, self.kwargs = code, kwargs

    def save(self, path):
        with open(path, 'w') as f:
            f.write(self.code)

    def execute(self, save_file_path=None):
        """Run the generated program (one process, or an mpiexec/torchrun fleet) and read back its step time."""
        keep = save_file_path is not None
        path = save_file_path or tempfile.NamedTemporaryFile(delete=False, suffix='.py').name
        with open(path, 'w') as f:
            f.write(self.code)
        log = tempfile.NamedTemporaryFile(delete=False, suffix='.log').name
        os.unlink(log)
        os.environ['CONFIG_STORE_PATH'] = log
        cmd = session.backend.get_execute_cmd(self.kwargs['spmd_nodes'], self.kwargs['total_nodes'], self.kwargs['device_type'], path)
        result = {}
        try:
            logging.info('Executing: %s' % cmd)
            assert os.system(cmd) == 0, 'Failed to execute command: %s' % cmd
            with open(log) as f:
                result = json.loads(f.read().strip())
        except Exception as ex:  # noqa
            logging.warning('program execution failed: %s', ex)
            result = {}
        finally:
            for p in ([] if keep else [path]) + [log]:
                try:
                    os.unlink(p)
                except FileNotFoundError:
                    pass
        return result


class Config:
    VERSION = '0.1'

    def __init__(self, config):
        if isinstance(config, str):
            with open(config) as f:
                config = json.load(f)
        if not isinstance(config, dict):
            raise Exception('Unsupported config value: %s' % config)
        if config['v'] != Config.VERSION:
            raise Exception('Incompatible config version: expect %s, got %s' % (Config.VERSION, config['v']))
        self.config = config

    @staticmethod
    def load_from_file(filename):
        return Config(filename) if filename is not None and os.path.exists(filename) else None

    @staticmethod
    def create(config, environ, timecost=0):
        return Config({'v': Config.VERSION, 't': timecost, 'b': config, 'kwargs': environ})

    def save(self, filepath):
        with open(filepath, 'w') as f:
            json.dump(self.config, f)

    def __str__(self):
        return json.dumps(self.config)


_UNSET = object()


def environ_config(kwargs):
    kwargs.setdefault('spmd_nodes', kwargs['total_nodes'])
    kwargs.setdefault('device_type', os.environ.get('DEVICE', 'cuda'))
    kwargs.setdefault('run_mode', os.environ.get('MODE', 'train'))
    assert kwargs['total_nodes'] % kwargs['spmd_nodes'] == 0, '`total_nodes` must be exactly divided by `spmd_nodes`.'
    return kwargs


class Custom:
    """A graph node: a data/param tensor (``data`` is a dict) or a compute op (``data`` is a layout string)."""

    def __init__(self, data, fw_ops=None, inputs=None, op_name=None, shape_fn=None, flops=None, depends=()):
        kind = op_name or inspect.currentframe().f_back.f_code.co_name
        if not re.match('^[a-zA-Z0-9]+$', kind):
            kind = 'Custom'
        assert kind[0].isupper(), 'The leading charactor of the operator name must be uppercase letter (received: "%s").' % kind
        bucket = 'id' if kind == 'Id' else ('builtin' if kind == 'Builtin' else 'op')
        serial = session.counters[bucket]
        session.counters[bucket] += 1
        self.name = '%s%s%d' % (kind[0].lower(), kind[1:], serial)
        self.depends = list(depends) if isinstance(depends, (list, tuple)) else [depends]
        self.outputs, self.data = [], data
        self.fw_ops = fw_ops.replace('@@', '') if fw_ops is not None else None

        if isinstance(data, dict):
            self.op_type = 'param' if data['is_param'] else 'data'
            if data['is_param']:
                self.name += '_'
            self.inputs, self.shape, self.dtype, self.flops = [], list(data['shape']), data['dtype'], flops or 0
        else:
            if inputs is None:
                assert fw_ops is not None, 'At least one property in "fw_ops" and inputs should be specified.'
                names = []
                for piece in fw_ops.split('@@')[1::2]:
                    if piece not in names:
                        names.append(piece)
                inputs = [session.custom_dict[n] for n in names]
            self.op_type, self.inputs, self.parser = 'compute', list(inputs), Layout(data)
            if shape_fn is not None:
                self.shape, self.dtype = shape_fn(self.inputs)
            else:
                try:
                    self.shape = self.parser.infer_shape([i.shape for i in self.inputs])
                except KeyError:
                    raise Exception('Cannot auto-infershape for op %s due to unknown dimension size by tensor format: %s' % (self.name, data))
                self.dtype = self.inputs[0].dtype
            self.flops = flops if flops is not None else self.parser.flops(self.shape, [i.shape for i in self.inputs])
        assert self.name not in session.custom_dict, 'Node with name `%s` has already existed in current session.' % self.name
        session.custom_dict[self.name] = self

    def __str__(self):
        return '@@%s@@' % self.name

    def numel(self):
        return product(self.shape)

    def get_input_by_name(self, name):
        for inp in self.inputs:
            if inp.name == name:
                return inp
        raise Exception('Node input with name `%s` not found!' % name)

    # ---- graph analysis ---------------------------------------------------------------------------------------
    def _link_consumers(self, parent, kwargs):
        if parent is not None and parent not in self.outputs:
            self.outputs.append(parent)
        # States pinned by an earlier serialisation (e.g. one with spmd_nodes == 1) must not leak into the next one:
        # remember what the user pinned (strict format / node.config = ...) the first time and start from that.
        if '_user_config' not in self.__dict__:
            self._user_config = self.__dict__.get('config', _UNSET)
        if self._user_config is _UNSET:
            self.__dict__.pop('config', None)
        else:
            self.config = self._user_config
        if kwargs['spmd_nodes'] == 1:
            self.config = -1
        elif session.ptype == 'dp':
            self.config = -1 if self.op_type == 'param' else 0
        elif session.ptype == 'zero':
            self.config = -2 if self.op_type == 'param' else 0
        elif self.name in session.manual_config:
            self.config = session.manual_config[self.name]
        for inp in self.inputs:
            inp._link_consumers(self, kwargs)

    update_config = lambda self, parent, **kw: self._link_consumers(parent, kw)   # reference name

    def _topological(self):
        order, seen = [], set()

        def visit(n):
            seen.add(id(n))
            for i in n.inputs:
                if id(i) not in seen:
                    visit(i)
            order.append(n)
        visit(self)
        return order

    def articulare_analyse(self):
        """Cut the (undirected, parameter-free) graph at articulation points -> list of (stage ops, shared inputs)."""
        index, low, cut, stack = {}, {}, {}, []
        counter = [0]

        def dfs(u, root):
            counter[0] += 1
            index[u] = low[u] = counter[0]
            stack.append(u)
            children = 0
            for v in u.inputs + u.outputs:
                if v.op_type == 'param':
                    continue
                if v not in index:
                    children += 1
                    dfs(v, root)
                    low[u] = min(low[u], low[v])
                    if (u is root and children > 1) or (u is not root and low[v] >= index[u]):
                        cut[u] = cut.get(u, 0) + 1
                    if low[v] >= index[u]:
                        while stack.pop() is not v:
                            pass
                else:
                    low[u] = min(low[u], index[v])
            cut[u] = cut.get(u, 0) + 1
        dfs(self, self)

        stages, visited, serial = {}, set(), [0]

        def collect(u, sid, leader):
            if u in visited or u.op_type != 'compute':
                return
            stages.setdefault(sid, []) if leader else None
            stages[sid].append(u)
            visited.add(u)
            for v in u.inputs:
                if cut.get(v, 0) > 1:
                    serial[0] += 1
                    collect(v, serial[0], True)
                else:
                    collect(v, sid, False)
        collect(self, 0, True)

        groups = []
        for _, members in sorted(stages.items(), reverse=True):
            shared = set()
            for op in members:
                shared = {y for y in op.inputs if len(y.outputs) > 1}
            groups.append((list(reversed(members)), shared))
        return groups

    def serialize(self, **kwargs):
        self._link_consumers(None, kwargs)
        groups = self.articulare_analyse()
        order = self._topological()
        pinned = {n.name: n.config for n in order if hasattr(n, 'config')}
        inputs = [n for n in order if isinstance(n.data, dict)]
        computes = [n for n in order if not isinstance(n.data, dict)]
        return groups, computes, inputs, pinned

    # ---- user entry points --------------------------------------------------------------------------------------
    def get_data_parallel_config(self, **kwargs):
        cfg = {n.name: ([-1, ''] if n.op_type == 'param' else [0, 'BAR:0']) for n in self._topological()}
        return Config.create(cfg, environ_config(kwargs))

    def autotune(self, config_file=None, **kwargs):
        cached = Config.load_from_file(config_file)
        if cached:
            return cached
        kwargs, results = optimize(self, **kwargs)
        valid = [sol for _, sol in results if sol is not None]
        if not valid:
            raise Exception('No valid configuration found!')
        best_time, best = min(valid, key=lambda r: r[0])
        cfg = Config.create(best, kwargs, best_time)
        if config_file is not None:
            cfg.save(config_file)
        return cfg

    def compile(self, config, **kwargs):
        if not isinstance(config, dict):
            assert config.config['v'] == Config.VERSION
            config.config['kwargs'].update(kwargs)
            kwargs, config = config.config['kwargs'], config.config['b']
        total, group = kwargs['total_nodes'], kwargs['spmd_nodes']
        assert total % group == 0, '`total_nodes` must by evenly divided by `spmd_nodes`, got: %d %% %d != 0' % (total, group)
        _, computes, inputs, pinned = self.serialize(**kwargs)
        for n in computes + inputs:
            st = config[n.name][0]
            if pinned.get(n.name, st) != st:
                raise Exception('Unstatisfied sharding state requirements on node `%s`' % n.name)
            if st >= 0 and n.shape[st] % group != 0:
                raise Exception('Unstatisfied slicing chunks `%d // %d` on node `%s`' % (n.shape[st], group, n.name))

        be = session.backend
        input_defs, param_defs = [], []
        for n in inputs:
            code = be.get_input_definition(n.name, n.shape, config[n.name][0], n.dtype, is_param=(n.op_type == 'param'))
            (param_defs if n.op_type == 'param' else input_defs).append((n.name, code))

        body, temps = [], 0
        for n in computes:
            state, key = config[n.name]
            choice = None
            if ':' in key:
                key, choice = key.split(':')
                choice = int(choice)
            plans = []
            try:
                plans = list(solver.PATTERNS[key](session, n, state, group, choice))
            except NotImplementedError:
                pass
            assert len(plans) <= 1, 'Ambiguous solution `%s` for node with `%s` at dimension %s' % (key, n.name, state)
            assert plans, 'No statisfied parallel pattern `%s` applying on node `%s`' % (key, n.name)
            _, need, links = plans[0]
            line = '%s = %s' % (n.name, n.fw_ops)
            for i, inp in enumerate(n.inputs):
                expr, have, want = inp.name, config[inp.name][0], need[i]
                if have != want:
                    extra = {'output_shape': inp.shape, 'is_param': inp.op_type == 'param'}
                    if have == -2 and want >= 0:
                        expr = be.link(be.link(expr, -2, -1, **extra), -1, want, **extra)
                    else:
                        expr = be.link(expr, have, want, **extra)
                if i in links:
                    expr = links[i].replace('$', expr).strip() or expr
                if expr != inp.name:
                    temps += 1
                    line = '_temp%d = %s; ' % (temps, expr) + re.sub(r'\b%s\b' % re.escape(inp.name), '_temp%d' % temps, line)
            body.append(line)
            post = links.get('', '').replace('$', n.name).strip()
            if post:
                body.append('%s = %s' % (n.name, post))

        headers, seen = [], set()

        def emit(deps):
            for d in deps:
                if id(d) in seen:
                    continue
                seen.add(id(d))
                emit(d['depends'])
                headers.append(d['data'])
        for n in computes:
            emit(n.depends)
        code = be.generate_framework_code(kwargs['device_type'], group, total // group, kwargs['run_mode'], self.name,
                                          headers, input_defs, param_defs, body)
        return Program(code, kwargs)


def optimize(node, **kwargs):
    kwargs = environ_config(kwargs)
    if session.is_strict_fmt:
        node = Id(node, op_name='Builtin')
        node.config = 0
    groups, computes, inputs, pinned = node.serialize(**kwargs)
    print('<< TUNE Graph >>\n')
    for n in inputs:
        print('| %s <- new_%s() | %s%s | %s |' % (n.name, n.op_type, n.dtype, n.shape, getattr(n, 'config', None)))
    print('---------------------------------------------------')
    for n in computes:
        print('| %s <- %s | %s%s | "%s" | %s |' % (n.name, ', '.join(i.name for i in n.inputs), n.dtype, n.shape, n.data, getattr(n, 'config', None)))
    print('\n>> config = %s\n' % json.dumps(pinned))
    return kwargs, solver.solve_partition(session, groups, input_nodes=inputs, split_pref=pinned, kwargs=kwargs)


def Id(x, op_name=None):
    axes = ''.join(chr(ord('a') + i) for i in range(len(x.shape)))
    return Custom('%s = %s' % (axes, axes), '%s' % x, op_name=op_name)


def Tensor(shape, dtype, is_param=False):
    node = Custom({'shape': shape, 'dtype': dtype, 'is_param': is_param}, inputs=[])
    if not is_param and session.is_strict_fmt:
        pinned = getattr(node, 'config', session.manual_config.pop(node.name, None))
        node.config = 0
        node = Id(node, op_name='Builtin')
        if pinned is not None:
            node.config = pinned
    return node
