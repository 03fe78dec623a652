"""Run code written against ``tutel`` on top of tutel_b200 without editing it.

    import tutel_b200.compat as compat
    compat.install_as_tutel()          # afterwards: `from tutel import moe, net, system, jit` resolve to tutel_b200
"""
import importlib
import sys
import types

_ALIASES = {
    'tutel': 'tutel_b200',
    'tutel.moe': 'tutel_b200.moe',
    'tutel.net': 'tutel_b200.net',
    'tutel.system': 'tutel_b200.system',
    'tutel.jit': 'tutel_b200.jit',
    'tutel.impls': None,
    'tutel.impls.moe_layer': 'tutel_b200.models.moe_layer',
    'tutel.impls.fast_dispatch': 'tutel_b200.ops.dispatch',
    'tutel.impls.communicate': 'tutel_b200.parallel.communicate',
    'tutel.impls.overlap': 'tutel_b200.parallel.overlap',
    'tutel.impls.losses': 'tutel_b200.models.losses',
    'tutel.gates': None,
    'tutel.gates.top': 'tutel_b200.models.gates.top',
    'tutel.gates.cosine_top': 'tutel_b200.models.gates.cosine_top',
    'tutel.experts': None,
    'tutel.experts.ffn': 'tutel_b200.models.experts.ffn',
    'tutel.experts.llama_ffn': 'tutel_b200.models.experts.llama_ffn',
    'tutel.checkpoint': 'tutel_b200.checkpoint',
    'tutel.checkpoint.gather': 'tutel_b200.checkpoint.gather',
    'tutel.checkpoint.scatter': 'tutel_b200.checkpoint.scatter',
    'tutel.parted': 'tutel_b200.parted',
    'tutel.parted.spmdx': 'tutel_b200.parted.spmdx',
}


def install_as_tutel(force: bool = False) -> None:
    if 'tutel' in sys.modules and not force and getattr(sys.modules['tutel'], '__tutel_b200__', False):
        return
    for alias, target in _ALIASES.items():
        if target is None:
            mod = types.ModuleType(alias)
            mod.__path__ = []
        else:
            try:
                mod = importlib.import_module(target)
            except ImportError:
                continue
        sys.modules[alias] = mod
    sys.modules['tutel'].__tutel_b200__ = True
    # dispatch helpers that lived in fast_dispatch in the reference
    disp = sys.modules['tutel.impls.fast_dispatch']
    from .ops import routing
    for name in ('extract_critical', 'get_dispatch_count'):
        if not hasattr(disp, name):
            setattr(disp, name, getattr(routing, name))
    for parent, child in (('tutel.impls', 'moe_layer'), ('tutel.impls', 'fast_dispatch'), ('tutel.impls', 'communicate'),
                          ('tutel.impls', 'overlap'), ('tutel.impls', 'losses'), ('tutel.gates', 'top'),
                          ('tutel.gates', 'cosine_top'), ('tutel.experts', 'ffn'), ('tutel.experts', 'llama_ffn')):
        if parent in sys.modules and parent + '.' + child in sys.modules:
            setattr(sys.modules[parent], child, sys.modules[parent + '.' + child])
