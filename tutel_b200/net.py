"""Public collective API (mirrors tutel/net.py:6-12)."""
from .parallel.communicate import (get_world_size, get_world_rank, create_groups_from_world, create_standalone_group,
                                   barrier)
# communication without backward compute
from .parallel.communicate import (simple_all_reduce, simple_all_to_all, simple_split, simple_reduce_scatter,
                                   simple_all_gather)
# communication with backward compute
from .parallel.communicate import (all_to_all, all_to_all_single, all_gather, zero_gather, zero_scatter, spatial_split,
                                   reduce_scatter, allreduce_forward, allreduce_backward)
# ragged batch collectives
from .parallel.communicate import batch_all_to_all_v, batch_all_gather_v
from .parallel.optimizer import TutelDistributedOptimizer

__all__ = ['get_world_size', 'get_world_rank', 'create_groups_from_world', 'create_standalone_group', 'barrier',
           'simple_all_reduce', 'simple_all_to_all', 'simple_split', 'simple_reduce_scatter', 'simple_all_gather',
           'all_to_all', 'all_to_all_single', 'all_gather', 'zero_gather', 'zero_scatter', 'spatial_split',
           'reduce_scatter', 'allreduce_forward', 'allreduce_backward', 'batch_all_to_all_v', 'batch_all_gather_v',
           'TutelDistributedOptimizer']
