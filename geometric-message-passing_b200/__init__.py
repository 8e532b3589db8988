"""gmp_b200 -- B200-native geometric message passing (hot path of NW-JEFF/Geometric-Message-Passing).

Layout
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/gmp_b200.h -> libgmp_b200.so)
  _lib.py      ctypes binding of that ABI (no fallback: missing library = error)
  graph.py     radius_graph (torch_cluster canonical order), CSR / Graph cache
  scatter.py   torch_scatter.scatter replacements (deterministic segmented reductions)
  schnet.py    SchNet InteractionBlock / CFConv / SchNetModel     (models/schnet.py)
  egnn.py      EGNNLayer / MPNNLayer / EGNNModel                   (models/layers/egnn_layer.py, models/egnn.py)
  tfn.py       TensorProductConvLayer / TFNModel                   (models/layers/tfn_layer.py, models/tfn.py)
  mace.py      SymmetricContraction / EquivariantProductBasisBlock / MACEModel (models/mace_modules, models/mace.py)
  mace_blocks.py  the ACEsuit-style 'uvu' interaction blocks       (models/mace_modules/blocks.py:136-530)
  gvp.py       GVP / GVPConvLayer / GVPGNNModel                     (models/layers/gvp_layer.py, models/gvpgnn.py)
"""
from . import _lib  # noqa: F401
from .graph import CSR, Graph, build_csr, get_graph, radius_graph  # noqa: F401
from .scatter import scatter, scatter_mean, scatter_sum, segment_reduce  # noqa: F401
from .schnet import (CFConv, GaussianSmearing, InteractionBlock, SchNetModel, ShiftedSoftplus,  # noqa: F401
                     edge_length, global_add_pool, global_mean_pool)

from .egnn import EGNNLayer, EGNNModel, MPNNLayer  # noqa: F401
from .irreps import Irreps  # noqa: F401
from .tfn import (BatchNorm, Gate, RadialEmbeddingBlock, SphericalHarmonics, TensorProductConvLayer,  # noqa: F401
                  TensorProductPlan, TFNModel, edge_geometry, first_node_pooling)
from .mace import (Contraction, EquivariantLinear, EquivariantProductBasisBlock, MACEModel,  # noqa: F401
                   SymmetricContraction, reshape_irreps)
from . import mace_blocks  # noqa: F401,E402
from .mace_blocks import (AgnosticNonlinearInteractionBlock, AgnosticResidualNonlinearInteractionBlock,  # noqa: F401,E402
                          RealAgnosticInteractionBlock, RealAgnosticResidualInteractionBlock,
                          ResidualElementDependentInteractionBlock, UVUTensorProduct)
from . import gvp  # noqa: F401,E402
from .gvp import GVP, GVPConv, GVPConvLayer, GVPGNNModel  # noqa: F401,E402
from .data import Batch, Data, DataLoader, DevicePrefetcher, coalesce, to_undirected  # noqa: F401,E402
from .graphs import GraphedStep  # noqa: F401,E402
from . import distributed  # noqa: F401,E402
from .distributed import PartitionedEGNN, SlabPartition, allreduce_gradients, halo_exchange, slab_partition  # noqa: F401,E402


def set_fast_matmul(enabled: bool = True) -> None:
    """Node-side `nn.Linear` layers are plain library GEMMs (cuBLAS).  In precision="bf16" runs they may use TF32
    tensor cores (10-bit mantissa, well inside that mode's 1e-2 tolerance); the fp32-strict mode must keep this off.
    This flips PyTorch's process-wide matmul precision switch."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = bool(enabled)
    torch.backends.cudnn.allow_tf32 = bool(enabled)


def load_reference_checkpoint(module, state_dict: dict):
    """Load a `state_dict` saved from the reference's module of the same name into a gmp_b200 module.

    Every parameter and buffer of `module` must be present in `state_dict` (else KeyError: the checkpoint is for another
    architecture).  Keys the fused modules do not own are dropped and returned: the reference's TFN / MACE checkpoints
    carry constructor-time e3nn buffers (`tp.*`, `gate.*`, `linear.*` output masks) and the Bessel / cutoff constants
    (`radial_embedding.bessel_fn.{bessel_weights,r_max,prefactor}`, `cutoff_fn.{p,r_max}`) that are recomputed here from
    the constructor arguments, so a plain `load_state_dict(strict=True)` raises on them."""
    own = set(module.state_dict().keys())
    missing = sorted(own - set(state_dict))
    if missing:
        raise KeyError(f"load_reference_checkpoint: absent from the checkpoint: {missing}")
    module.load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=True)
    return sorted(set(state_dict) - own)


__version__ = "0.2.0"
