"""nwx -- B200-native (sm_100a) engine for the NeRF render/train hot path of
dmjovan/NeRF-Workspaces-Explorer, behind the reference's own Python entry points.

    from nwx import create_rays, sample_pdf, run_network, raw2outputs, NeRFModel, Embedding
    from nwx import NeRFReplicaInferenceHandler, COORD

Everything computes in hand-written CUDA kernels loaded through the C ABI of libnwx.so
(include/nwx.h); importing the package without the built library is fine, calling into it is
an error (there is no fallback path)."""
from . import config, dist, synthetic                                     # noqa: F401
from ._lib import NwxError, build, lib                                     # noqa: F401
from .batch_utils import batchify, batchify_rays                           # noqa: F401
from .camera_poses import get_camera_poses_from_list_of_coordinates        # noqa: F401
from .data_descriptors import COORD, HW, XYZ                               # noqa: F401
from .engine import COARSE, FINE, REFERENCE_KEYS, Engine                   # noqa: F401
from .inference import NeRFReplicaInferenceHandler                         # noqa: F401
from .models import (Embedding, NeRFModel, img2mse, mse2psnr, raw2outputs, run_network,   # noqa: F401
                     to8b, to8b_np)
from .rays import create_rays, sample_pdf                                  # noqa: F401
from .reference_patch import patch_reference, unpatch_reference            # noqa: F401
from .training import NeRFReplicaTrainingHandler, Trainer                  # noqa: F401

from .workspace import (OfficeBelgradeWorkspace, OfficeGeneveWorkspace, OfficeNewYorkWorkspace,   # noqa: F401
                        OfficeTokyoWorkspace, Workspace)

__version__ = "0.1.0"
