"""B200-native IHub/LHub neighbourhood link prediction: the package around the C ABI.

    csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/nlp_b200.h)
    binding.py ctypes mirror of the ABI (tests, bench.py)
    graphs.py  deterministic synthetic workloads (R-MAT, road lattice, web-crawl, planted partition)
    distributed.py  multi-GPU candidate all-gather + merge (torch.distributed: NCCL / gloo)
    build.py   nvcc recipe

The directory name carries the reference's name (with hyphens), so import it through the
repo-root shim ``nlp_b200``.
"""
from . import build, binding, graphs, distributed  # noqa: F401
from .binding import Predictor, MEASURES, UNBOUNDED, NlpError  # noqa: F401
