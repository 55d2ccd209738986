"""pumi-pic_b200: B200-native implementation of PUMI-PIC's particle hot path.

The product is libpumipic_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/pumipic_b200.h) plus the header-only C++ mirror of the reference API in cpp/.
This Python package only binds the C ABI for tests and bench.py: torch tensors provide device
memory and streams.  Import with importlib.import_module("pumi-pic_b200").
"""
from . import capi                                    # noqa: F401
from .capi import PumipicError, lib                  # noqa: F401
from .api import (Mesh, ParticleStructure, SearchResult, search_mesh, push_constant,  # noqa: F401
                  push_direction, update_positions, push_direction_search, host_kuhn_cube,
                  host_plate, host_derive_sides, push_from, elliptical_setup, elliptical_push,
                  set_unsafe_procs, gyro_ring_map, gyro_scatter, gyro_interleave, Comm, migrate, host_picpart_tags, host_picpart_tags_bridged,
                  host_entity_owners, push_direction_search_host, push_boris, gather_tet_field, gather_grid2d,
                  gather_grid2d_vector, gather_grid3d, host_picpart_extract, CommPlan, HostMesh, Picpart, host_read_partition,
                  FULL, BFS, MINIMUM, NONE, Balancer, host_lb_plan, trace_particle_through_mesh)
