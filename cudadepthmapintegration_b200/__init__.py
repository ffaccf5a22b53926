"""B200-native depth-map integration engine: the hot path of bastienjacquet/CudaDepthMapIntegration
(per-voxel TSDF integration + per-point mesh coloration) behind the reference's own operator API.
See DESIGN.md; the numeric work lives in csrc/ (hand-written sm_100a CUDA behind include/dmi_b200.h)."""
from .engine import Context, DmiError, Group  # noqa: F401
from .reconstruction import CudaReconstructionFilter, cuda_initialize, process_depth_map  # noqa: F401
from .coloration import MeshColoration  # noqa: F401
