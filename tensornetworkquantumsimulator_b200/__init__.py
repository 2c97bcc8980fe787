"""tnqs-cuda: B200-native engine for the BP message sweep and the two-site gate application of
TensorNetworkQuantumSimulator, behind the reference's API names.  The compute lives in
`csrc/` (libtnqs_cuda.so, C-ABI in include/tnqs_cuda.h); this package is the host-side mirror of the
reference interface for that path and never falls back to a CPU implementation."""
from .graphs import (Graph, named_grid, named_comb_tree, path_graph, named_hexagonal_lattice_graph,
                     heavy_hexagonal_lattice, eagle_heavy_hex, build_graph_from_gates, build_graph_from_circuit,
                     forest_cover_edge_sequence, edge_color, grid_edge_color)
from . import gates, rng
