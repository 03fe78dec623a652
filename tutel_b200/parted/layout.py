"""Einsum-like layout strings: ``"mn += mk, kn"``, grouped axes ``"(ab)c = abc"``, wildcard ``*``.

The character before ``=`` selects the reduction of axes that vanish from the output (``+`` sum, ``<``/``>`` min/max,
``[``/``]`` arg-variants); only ``+`` reductions can be distributed with an all-reduce / reduce-scatter.
(Reference behaviour: tutel/parted/spmdx.py:45-130.)
"""
from typing import Dict, List, Tuple, Union

Axis = Union[str, List[str]]
REPLICATED, ZERO_SHARDED = -1, -2


def _tokenize(spec: str) -> List[Axis]:
    out: List[Axis] = []
    group = None
    for ch in spec:
        if ch.isspace():
            continue
        if ch == '(':
            group = []
        elif ch == ')':
            out.append(group if len(group) > 1 else group[0])
            group = None
        elif group is not None:
            group.append(ch)
        else:
            out.append(ch)
    return out


class Operand:
    """Axes of one tensor: ``axes[i]`` is a name, ``'*'`` or a list of fused names; ``where[name]`` -> dim or (dim, pos)."""

    def __init__(self, spec: str):
        self.axes = _tokenize(spec)
        self.where: Dict[str, Union[int, Tuple[int, int]]] = {}
        for dim, ax in enumerate(self.axes):
            if isinstance(ax, str):
                if ax != '*':
                    self.where[ax] = dim
            else:
                for pos, name in enumerate(ax):
                    self.where[name] = (dim, pos)


class Layout:
    def __init__(self, text: str):
        lhs, rhs = text.split('=')
        lhs = lhs.rstrip()
        self.reduce_type = ''
        if lhs and lhs[-1] in '+<>[]':
            self.reduce_type, lhs = lhs[-1], lhs[:-1]
        self.out = Operand(lhs)
        self.ins = [Operand(part) for part in rhs.split(',')]

    @property
    def num_inputs(self) -> int:
        return len(self.ins)

    def reduce_axes(self) -> List[str]:
        seen, order = set(), []
        for op in self.ins:
            for name in op.where:
                if name not in self.out.where and name not in seen:
                    seen.add(name)
                    order.append(name)
        return order

    def sources_of_axis(self, name: str):
        """Which dim of every input carries axis ``name`` (-1: that input does not have it)."""
        if name == '*':
            raise NotImplementedError()
        if not isinstance(name, str):
            name = name[0]          # a fused output dim is split along its leading axis
        dims, parted = {}, 0
        for i, op in enumerate(self.ins):
            pos = op.where.get(name)
            if pos is None:
                dims[i] = REPLICATED
                continue
            if isinstance(pos, tuple):
                if pos[1] != 0:
                    raise NotImplementedError()   # only the leading member of a fused dim can be split
                pos = pos[0]
            dims[i] = pos
            parted += 1
        return dims, parted

    def sources_of_output_dim(self, dim: int):
        if dim == REPLICATED:
            return {i: REPLICATED for i in range(self.num_inputs)}, 0
        if dim < 0 or self.out.axes[dim] == '*':
            raise NotImplementedError()
        return self.sources_of_axis(self.out.axes[dim])

    def infer_shape(self, input_shapes) -> List[int]:
        size = {}
        for op, shape in zip(self.ins, input_shapes):
            for name, pos in op.where.items():
                if isinstance(pos, int):
                    size[name] = shape[pos]
        shape = []
        for ax in self.out.axes:
            if isinstance(ax, str):
                shape.append(size[ax])
            else:
                n = 1
                for a in ax:
                    n *= size[a]
                shape.append(n)
        return shape

    def flops(self, out_shape, input_shapes) -> int:
        n = 1
        for v in out_shape:
            n *= int(v)
        if self.reduce_type:
            size = {}
            for op, shape in zip(self.ins, input_shapes):
                for name, pos in op.where.items():
                    if isinstance(pos, int):
                        size[name] = shape[pos]
            for name in self.reduce_axes():
                n *= int(size.get(name, 1))
            n *= 2
        return n
