"""vivae_b200 -- B200-native (sm_100a) implementation of the ViVAE hot path.

Drop-in for the reference package's public surface on that path
(/root/reference/src/vivae/__init__.py:31-53):

    import vivae_b200 as vv
    knn = vv.make_knn(x, k=100)              # exact kNN graph, tensor cores + FP32 re-rank
    x_d = vv.smooth(x, knn, k=50, coef=1.)   # Gauss-Seidel kNN smoother (bit-exact vs reference)
    model = vv.ViVAE(input_dim=x.shape[1]); model.fit(x_d); emb = model.transform(x_d)

Environment flags keep the reference's names (__init__.py:13-15):
  VIVAE_DETERMINISTIC (default "1")  -> torch.use_deterministic_algorithms
  VIVAE_CUDA / VIVAE_MPS             -> accepted and ignored: the backend switch of the
                                        reference collapses to CUDA; there is no CPU fallback.
Unlike the reference, importing this package does NOT call `torch.set_default_device`;
tensors created by user code stay where the user puts them.
"""

import logging
import os

DETERMINISTIC = os.environ.get("VIVAE_DETERMINISTIC", default="1") != "0"
if DETERMINISTIC:
    # cuBLAS needs this for deterministic GEMMs once use_deterministic_algorithms(True) is set
    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")

import torch  # noqa: E402

logger = logging.getLogger(__name__)

if torch.cuda.is_available():
    DEVICE = torch.device("cuda", torch.cuda.current_device())
    DEVICE_NAME = "cuda"
else:  # import still works (host-side logic, CPU tests); every compute entry point raises
    DEVICE = torch.device("cpu")
    DEVICE_NAME = "cpu"

torch.use_deterministic_algorithms(DETERMINISTIC)

from . import _lib  # noqa: E402
from .knn import correct_knn_for_duplicates, ensure_valid_knn, make_knn, smooth  # noqa: E402
from .losses import MDSLoss, quartet_loss  # noqa: E402
from .model import ViVAE  # noqa: E402
from .network import Autoencoder  # noqa: E402

logger.info('initialised with "%s" backend and determinism %s', DEVICE_NAME, "enabled" if DETERMINISTIC else "disabled")

__all__ = [
    "DEVICE",
    "DEVICE_NAME",
    "DETERMINISTIC",
    "ViVAE",
    "Autoencoder",
    "MDSLoss",
    "quartet_loss",
    "make_knn",
    "smooth",
    "ensure_valid_knn",
    "correct_knn_for_duplicates",
]
__version__ = "0.0.2+b200.1"
