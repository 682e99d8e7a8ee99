"""Same re-exports as the reference ``Utils/__init__.py:1-2``."""
from solidspy_opt.Utils.beams import beam, beam_1, beam_2, rect_grid
from solidspy_opt.Utils.solver import (protect_els, del_node, volume, sensitivity_elsBESO, adjacency_nodes,
                                       center_els, sensitivity_nodes, sensitivity_filter, sensitivity_elsESO,
                                       strain_els, protect_elsESO, del_nodeESO, sparse_assem,
                                       optimality_criteria, density_filter)
