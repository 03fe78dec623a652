"""Measured dynamic programming over sharding states (reference: tutel/parted/solver.py:19-144).

The graph is cut at articulation points into *stages*; inside a stage every op is visited in topological order and,
for each admissible output state (replicated or split on a dim), every pattern that can produce it from already-solved
input states is compiled into a runnable program and timed.  ``best[node][state] = (time, {node: (state, pattern)})``.
Tensors consumed by several ops of a stage are pinned to one state per enumeration pass so that all consumers agree.
"""
import itertools
import json
import sys

from .layout import REPLICATED, ZERO_SHARDED
from .patterns import PATTERNS, register_primitive  # noqa: F401  (re-export)

spmd_primitives_dict = PATTERNS


def _merge(into, record):
    """Union two {node: (state, pattern)} assignments; None on conflict."""
    if record is None:
        return None
    for k, v in record[1].items():
        if into.setdefault(k, v) != v:
            return None
    return into


def _input_seed_states(node, group_size, pinned):
    if node.name in pinned:
        return [pinned[node.name]]
    states = [REPLICATED, ZERO_SHARDED]
    states += [d for d in range(len(node.shape)) if node.shape[d] % group_size == 0]
    return states


def solve_partition(sess, compute_groups, input_nodes, split_pref, kwargs):
    group_size, total = kwargs['spmd_nodes'], kwargs['total_nodes']
    print('\nDistributed for total_nodes = %d, spmd_nodes = %d, run_mode = `%s`\n' % (total, group_size, kwargs['run_mode']))
    final_node = compute_groups[-1][0][-1]

    best = {}
    for inp in input_nodes:
        best[inp.name] = {st: (0.0, {inp.name: (st, '')}) for st in _input_seed_states(inp, group_size, split_pref)}

    for stage_nodes, shared in compute_groups:
        shared = list(shared)
        choices = [range(-1, len(n.shape)) for n in shared]
        stage_best = {}
        passes = list(itertools.product(*choices)) if shared else [()]
        for pass_id, pinned_states in enumerate(passes):
            pinned = {n.name: st for n, st in zip(shared, pinned_states)}
            for node in stage_nodes:
                table = best[node.name] = {}
                if group_size == 1:
                    candidates = [REPLICATED]
                elif node.name in split_pref:
                    candidates = [split_pref[node.name]]
                else:
                    candidates = range(-1, len(node.shape))
                for state in candidates:
                    if pinned.get(node.name, state) != state:
                        continue
                    if state >= 0 and node.shape[state] % group_size != 0:
                        continue
                    programs = []
                    for key, rule in PATTERNS.items():
                        assignment = None
                        try:
                            for choice, in_states, _ in rule(sess, node, state, group_size, None):
                                trial = {node.name: (state, '%s:%s' % (key, choice))}
                                for idx, st in in_states.items():
                                    src = node.inputs[idx].name
                                    rec = best.get(src, {}).get(st) if best.get(src) else None
                                    if rec is None or pinned.get(src, st) != st:
                                        trial = None
                                        break
                                    trial = _merge(trial, rec)
                                    if trial is None:
                                        break
                                if trial:
                                    assignment = trial
                                    break
                            if assignment:
                                prog = node.compile(assignment, **kwargs)
                                if prog:
                                    programs.append((prog, assignment))
                        except NotImplementedError:
                            continue
                    winner = (float('inf'), None)
                    for i, (prog, cfg) in enumerate(programs):
                        print('>> Try `%s:%s [ENUM:%d/%d]` (%d/%d), config = %s' % (node.name, state, pass_id + 1, len(passes), i + 1, len(programs), json.dumps(cfg)))
                        if len(passes) == 1 and len(programs) == 1 and node.name != final_node.name:
                            cost = -1          # nothing to choose: skip the measurement
                        else:
                            print('>> Program Snapshot:')
                            print(prog.code)
                            cost = prog.execute().get('step_time', float('inf'))
                        if cost < winner[0]:
                            winner = (cost, cfg)
                    if winner[1] is not None:
                        table[state] = winner
                        print('>> FL_%s_%s [ENUM:%d/%d] = %s\n' % (node.name, state, pass_id + 1, len(passes), winner))
            for state, rec in best[stage_nodes[-1].name].items():
                if state not in stage_best or rec[0] < stage_best[state][0]:
                    stage_best[state] = rec
        for node in stage_nodes:
            best[node.name] = None          # interior results are folded into the stage output
        best[stage_nodes[-1].name] = stage_best
        print('>> Stage `%s` solved; valid output states: %s' % (stage_nodes[-1].name, sorted(stage_best)))
        sys.stdout.flush()

    out = best[final_node.name]
    return [(dim, out.get(dim)) for dim in range(-1, len(final_node.shape))]
