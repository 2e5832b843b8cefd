"""edge_diffusion_tts_b200 -- B200 (sm_100a) implementation of the few-step
sampling path of Krabbens/edge-diffusion-tts behind the reference's own API.

    from edge_diffusion_tts_b200 import CFG, DiffusionSchedule, EdgeDiffusionDecoder, EdgeInference

Host code is Python/PyTorch (device memory, streams, torch.distributed); all
arithmetic on the path is hand-written CUDA in ``lib/libedtts.so`` behind the C
ABI of ``include/edtts.h``.  No CPU fallback, no Triton, no torch.compile.
"""
from .config import CFG, get_device, set_seed
from .schedule import DiffusionSchedule, DPMSolverPP
from .vq import VectorQuantizer
from .fsq import FSQ, FSQEncoder
from .decoder import EdgeDiffusionDecoder
from .encoder import SemanticEncoder
from .inference import EdgeInference
from .conv import DepthwiseSeparableConv
from .audio import normalize_mel, denormalize_mel, InverseMelScale, GriffinLim
from .longform import MelStitcher, chunk_plan, crossfade_window, generate_longform
from . import dist

__version__ = "0.1.0"
__all__ = ["CFG", "get_device", "set_seed", "DiffusionSchedule", "DPMSolverPP", "VectorQuantizer", "FSQ", "FSQEncoder", "EdgeDiffusionDecoder",
           "SemanticEncoder", "EdgeInference", "DepthwiseSeparableConv", "normalize_mel", "denormalize_mel", "InverseMelScale", "GriffinLim", "MelStitcher",
           "chunk_plan", "crossfade_window", "generate_longform", "dist"]
