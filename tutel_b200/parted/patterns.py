"""The library of parallel patterns the solver may assign to an op (reference: tutel/parted/patterns.py:12-129).

A pattern is a generator ``fn(session, node, out_state, group_size, choice)`` yielding
``(choice_index, input_states, connectors)``:

* ``out_state``      sharding state wanted for the op's output (-1 replicated, >= 0 split on that dim);
* ``input_states``   state every input must be in for the local compute to be valid;
* ``connectors``     code templates (``$`` = operand) applied to inputs (key = input index) and to the output (key '').

    BAR   pass-through: output split on a dim that is a batch-like axis of the inputs (no forward collective)
    FAR   inputs split on a contracted axis, forward all-reduce restores a replicated output
    RS    same, but reduce-scatter leaves the output split
    SPLIT replicated compute, keep only this rank's slice (spatial split)
    AG    compute on a slice, all-gather the output
    A2A   compute split on one dim, all-to-all re-partitions to another
    ZERO  parameters stored flat-sharded (ZeRO) and gathered just in time
"""
from .layout import REPLICATED, ZERO_SHARDED

PATTERNS = {}


def register_primitive(name=None):
    def deco(fn):
        key = name or 'custom_%d' % len(PATTERNS)
        assert key not in PATTERNS, 'Parallel Pattern with name `%s` already exists.' % key
        PATTERNS[key] = fn
        return fn
    return deco


def _replicated_input_links(sess, node, states):
    """Replicated operands of a sharded compute need their gradient all-reduced in backward."""
    return {i: sess.backend.link('$', REPLICATED, None, is_param=(node.inputs[i].op_type == 'param'))
            for i, st in states.items() if st == REPLICATED}


@register_primitive('BAR')
def pass_through(sess, node, out_state, group_size, choice):
    if out_state < REPLICATED:
        return
    states, parted = node.parser.sources_of_output_dim(out_state)
    if out_state == REPLICATED and parted == 0:
        yield 0, states, {}
        return
    yield 0, states, _replicated_input_links(sess, node, states)


def _contracted(sess, node, group_size, choice, output_link):
    if node.parser.reduce_type != '+':
        return
    for i, axis in enumerate(node.parser.reduce_axes()):
        if choice is not None and i != choice:
            continue
        try:
            states, parted = node.parser.sources_of_axis(axis)
        except NotImplementedError:
            continue
        assert parted > 0, 'It is unexpected that no certain input is parted.'
        links = _replicated_input_links(sess, node, states)
        links[''] = output_link
        yield i, states, links


@register_primitive('FAR')
def forward_allreduce(sess, node, out_state, group_size, choice):
    if out_state != REPLICATED:
        return
    yield from _contracted(sess, node, group_size, choice, sess.backend.link('$', None, REPLICATED))


@register_primitive('RS')
def forward_reduce_scatter(sess, node, out_state, group_size, choice):
    if out_state < 0:
        return
    yield from _contracted(sess, node, group_size, choice, sess.backend.link('$', None, out_state))


@register_primitive('SPLIT')
def spatial_split(sess, node, out_state, group_size, choice):
    if out_state < 0:
        return
    states, parted = node.parser.sources_of_output_dim(REPLICATED)
    assert parted == 0
    yield 0, states, {'': sess.backend.link('$', REPLICATED, out_state)}


@register_primitive('AG')
def forward_all_gather(sess, node, out_state, group_size, choice):
    if out_state != REPLICATED:
        return
    for dim in range(len(node.shape)):
        if choice is not None and dim != choice:
            continue
        if node.shape[dim] % group_size != 0:
            continue
        try:
            states, parted = node.parser.sources_of_output_dim(dim)
        except NotImplementedError:
            continue
        if parted == 0:
            continue
        links = _replicated_input_links(sess, node, states)
        links[''] = sess.backend.link('$', dim, REPLICATED)
        yield dim, states, links


@register_primitive('A2A')
def all_to_all(sess, node, out_state, group_size, choice):
    if out_state < 0:
        return
    shape = node.shape
    if len(shape) < 2 or shape[out_state] % group_size != 0:
        return
    for dim in range(len(shape)):
        if choice is not None and dim != choice:
            continue
        if dim == out_state or shape[dim] % group_size != 0:
            continue
        try:
            states, _ = node.parser.sources_of_output_dim(dim)
        except NotImplementedError:
            continue
        links = _replicated_input_links(sess, node, states)
        links[''] = sess.backend.link('$', dim, out_state)
        yield dim, states, links


@register_primitive('ZERO')
def zero_sharded_params(sess, node, out_state, group_size, choice):
    if out_state < 0:
        return
    states, parted = node.parser.sources_of_output_dim(out_state)
    if parted == 0:
        return
    links, any_param = {}, False
    for i, st in list(states.items()):
        if st != REPLICATED:
            continue
        if node.inputs[i].op_type == 'param':
            states[i] = ZERO_SHARDED
            links[i] = sess.backend.link('$', ZERO_SHARDED, REPLICATED, output_shape=node.inputs[i].shape)
            any_param = True
        else:
            links[i] = sess.backend.link('$', REPLICATED, None, is_param=False)
    if any_param:
        yield 0, states, links
