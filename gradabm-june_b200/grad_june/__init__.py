"""grad_june — B200-native drop-in for GradABM-JUNE's per-timestep infection path.

Same public names as the reference package (grad_june/__init__.py:1-9); the compute runs in
``libgradjune_b200.so`` (hand-written sm_100a CUDA behind a C ABI, see include/gradjune_b200.h).
"""
from .infection_networks import InfectionNetworks
from .transmission import TransmissionUpdater, TransmissionSampler
from .symptoms import SymptomsUpdater
from .infection import IsInfectedSampler
from .model import GradJune
from .timer import Timer
from .policies import Policies
from .runner import Runner
from .world import (HeteroData, ToUndirected, create_simple_connected_graph, layout_order_of, load_world,
                    make_synthetic_world, original_order, renumber_world)

__version__ = "0.1.0"
